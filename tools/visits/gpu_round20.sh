#!/bin/bash
# GPU-box visit 20 (validation of the final default build: 32-bit single-part probes on): GPU test suite, SF100 latencies +
# bytes-moved counters for all 22 queries, the same SF100 kernels without the 32-bit probes (A/B), bench line
set -u
mkdir -p gpurun_out
echo "== tests" ; timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -2 gpurun_out/tests.log
echo "== SF100, one GPU: latencies + counters"
Q1="q5,q17,q19,q3,q7,q10,q8,q20,q12,q13,q4,q18,q21,q1,q6,q9,q14,q15,q16,q2,q11,q22"
timeout 330 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --stats-so gpurun_variants/stats.so --queries $Q1 \
    --out gpurun_out/sf100_n1_all22_v6.json > gpurun_out/sf100_n1_all22_v6.log 2> gpurun_out/sf100_n1_all22_v6.err; echo "rc=$?"
python tools/show_tpch.py gpurun_out/sf100_n1_all22_v6.json profiles/r01_tpch_sf100_n1_all22_v5.json 2>/dev/null | tail -24 || grep -o '"query": "[a-z0-9]*", "sf": 100.0, "device_ms_min": [0-9.]*' gpurun_out/sf100_n1_all22_v6.log
echo "== the same without 32-bit probes"
SDQLB200_SO=gpurun_variants/noprobe32.so timeout 120 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --queries q5,q17,q19,q3,q7 --out gpurun_out/sf100_noprobe32.json > gpurun_out/sf100_noprobe32.log 2>&1; echo "rc=$?"
grep -o '"query": "[a-z0-9]*", "sf": 100.0, "device_ms_min": [0-9.]*' gpurun_out/sf100_noprobe32.log
echo "== bench" ; timeout 300 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_q1.json
du -sh gpurun_out
