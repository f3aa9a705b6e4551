#!/bin/bash
# 2-GPU visit: all 22 queries at SF2 on 2 ranks checked against the reference module, then the weak-scaling bench line
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== all 22 on 2 GPUs, SF2, checked"
timeout 900 $TR tools/run_tpch_dist.py --sf 2 --queries q1,q2,q3,q4,q5,q6,q7,q8,q9,q10,q11,q12,q13,q14,q15,q16,q17,q18,q19,q20,q21,q22 --check ref --reps 2 --out gpurun_out/dist2_sf2.json > gpurun_out/dist2_sf2.log 2>&1; echo rc=$?
python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/dist2_sf2.json")):
        print(r["query"], r.get("parity_vs_ref", r.get("error", "?"))[:80], "%.3f ms dev" % r.get("device_ms", -1), "%.2f ms wall" % r.get("latency_ms_wall", -1), "merges", r.get("merges_total"))
except Exception as e:
    print("no report:", e)
PY
tail -5 gpurun_out/dist2_sf2.log | cut -c1-300
echo "== bench N=2"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_q1_n2.json 2> gpurun_out/bench_q1_n2.err; echo rc=$?; cut -c1-1800 gpurun_out/bench_q1_n2.json; tail -3 gpurun_out/bench_q1_n2.err | cut -c1-300
