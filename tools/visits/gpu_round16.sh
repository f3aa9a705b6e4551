#!/bin/bash
# GPU-box visit 16: A/B of the per-iteration warp re-convergence (default) against the previous loops on identical data,
# all 22 queries at SF10 (results of the two builds compared); SF100 latencies + step traces with the new default;
# the device-side .tbl reader against the reference fixture
set -u
mkdir -p gpurun_out
echo "== .tbl reader"; timeout 300 python -m pytest tests/test_tbl.py -m gpu -q > gpurun_out/tests_tbl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests_tbl.log
echo "== A/B re-convergence, SF10"
timeout 500 python tools/ab_variants.py --sf 10 --reps 5 --variants default,noreconv --queries q12,q13,q9,q18,q21,q3,q4,q5,q7,q8,q10,q1,q6,q2,q11,q14,q15,q16,q17,q19,q20,q22 --out gpurun_out/ab16.json > gpurun_out/ab16.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/ab16.json"))
except Exception:
    r = [json.loads(l) for l in open("gpurun_out/ab16.log") if l.startswith("{")]
qs = []
for x in r:
    if x["query"] not in qs: qs.append(x["query"])
for q in qs:
    print(q, "  ".join("%s %.3f (%s)" % (x["variant"], x["device_ms_min"], x["vs_first_variant"][:12]) for x in r if x["query"] == q))
PY
echo "== SF100, one GPU"
Q1="q12,q13,q4,q18,q21,q1,q6,q9,q8,q3,q7,q10,q5,q14,q15,q17,q19,q20,q16,q2,q11,q22"
timeout 480 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --trace --queries $Q1 --out gpurun_out/sf100_n1_all22_v5.json > gpurun_out/sf100_n1_all22_v5.log 2> gpurun_out/sf100_n1_all22_v5.err; echo "rc=$?"
python tools/show_tpch.py gpurun_out/sf100_n1_all22_v5.json profiles/r01_tpch_sf100_n1_all22_v4_traced.json 2>/dev/null | tail -24
du -sh gpurun_out
