#!/bin/bash
# round 2, visit 5 (1 GPU): new bench.py (SF1 smoke, then the SF100 default), strided-dense key packing A/B at SF100,
# reference fingerprints at SF100 (BASELINE-named queries) and SF10 (all 22), ncu of q1_k0 at SF100
set -u
mkdir -p gpurun_out
echo "== bench smoke (SF1)"; timeout 300 python bench.py --sf 1 --steps 5 --queries q1,q3,q13 > gpurun_out/r02_bench_sf1.json 2> gpurun_out/r02_bench_sf1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r02_bench_sf1.json; tail -5 gpurun_out/r02_bench_sf1.err
echo "== bench SF100"; ( time timeout 900 python bench.py ) > gpurun_out/r02_bench_sf100_n1.json 2> gpurun_out/r02_bench_sf100_n1.err; echo "rc=$?"; cut -c1-3000 gpurun_out/r02_bench_sf100_n1.json; tail -8 gpurun_out/r02_bench_sf100_n1.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_bench_sf100_ref.json 2> gpurun_out/r02_bench_sf100_ref.err; echo "rc=$?"; cut -c1-1200 gpurun_out/r02_bench_sf100_ref.json; tail -5 gpurun_out/r02_bench_sf100_ref.err
echo "== stride A/B SF100"
Q="q3,q4,q5,q10,q12,q18,q21,q7,q8"
SDQLB200_STRIDE=0 timeout 300 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --queries $Q --out gpurun_out/r02_sf100_nostride.json > gpurun_out/r02_sf100_nostride.log 2>&1; echo "rc=$?"
timeout 300 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --queries $Q --out gpurun_out/r02_sf100_stride.json > gpurun_out/r02_sf100_stride.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
def load(f):
    try: return {x["query"]: x for x in json.load(open(f))}
    except Exception: return {x["query"]: x for x in (json.loads(l) for l in open(f.replace(".json", ".log")) if l.startswith("{"))}
a, b = load("gpurun_out/r02_sf100_nostride.json"), load("gpurun_out/r02_sf100_stride.json")
for q in a:
    if q in b: print(q, "nostride %.3f  stride %.3f  ws %.0f -> %.0f MB" % (a[q]["device_ms_min"], b[q]["device_ms_min"], a[q]["workspace_MB"], b[q]["workspace_MB"]))
PY
echo "== fingerprints SF100"; timeout 900 python tools/make_fingerprints.py --sf 100 --queries q1,q6,q3,q5,q9,q18 --out gpurun_out/tpch_sf100_fingerprints.json --report gpurun_out/r02_parity_sf100.json > gpurun_out/r02_parity_sf100.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_parity_sf100.log | cut -c1-300
echo "== fingerprints SF10"; timeout 900 python tools/make_fingerprints.py --sf 10 --out gpurun_out/tpch_sf10_fingerprints.json --report gpurun_out/r02_parity_sf10.json > gpurun_out/r02_parity_sf10.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_parity_sf10.log | cut -c1-200
echo "== ncu q1_k0 SF100"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:^q1_k0" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_q1_k0_sf100 -f python tools/run_tpch.py --sf 100 --device-gen --queries q1 --reps 1 > gpurun_out/ncu_r02_q1_k0_sf100.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_q1_k0_sf100.ncu-rep > gpurun_out/r02_q1_k0_sf100_ncu.txt 2>&1
ncu -i gpurun_out/r02_q1_k0_sf100.ncu-rep --page source --csv > gpurun_out/r02_q1_k0_sf100_source.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/r02_q1_k0_sf100_source.csv 40 > gpurun_out/r02_q1_k0_sf100_hot.txt 2>&1
rm -f gpurun_out/r02_q1_k0_sf100.ncu-rep gpurun_out/r02_q1_k0_sf100_source.csv
head -24 gpurun_out/r02_q1_k0_sf100_ncu.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --sf 100 --steps 2 --warmup 1 --queries none --no-e2e > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out
