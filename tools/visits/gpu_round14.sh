#!/bin/bash
# GPU-box visit 14 (short): SF100 latencies with the warp text scan, per-step driver traces, bytes-moved counters from the
# counting build; A/B of the text scan at SF10; SF10 parity of the changed paths against the reference module
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
Q1="q12,q13,q4,q18,q21,q1,q6,q9,q8,q3,q7,q10,q5,q14,q15,q17,q19,q20,q16,q2,q11,q22"
echo "== SF100, one GPU: latencies + traces + counters"
timeout 540 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --trace --stats-so gpurun_variants/stats.so --queries $Q1 \
    --out gpurun_out/sf100_n1_all22.json > gpurun_out/sf100_n1_all22.log 2> gpurun_out/sf100_n1_all22.err; echo "rc=$?"
tail -2 gpurun_out/sf100_n1_all22.err
echo "== A/B text scan, SF10"
timeout 200 python tools/ab_variants.py --sf 10 --reps 5 --queries q13,q9,q16 --variants default,notextscan --out gpurun_out/ab14.json > gpurun_out/ab14.log 2>&1; echo "rc=$?"
grep -o '"query": "[a-z0-9]*", "variant": "[a-z]*", "sf": 10.0, "device_ms_min": [0-9.]*' gpurun_out/ab14.log
grep -o '"vs_first_variant": "[^"]*"' gpurun_out/ab14.log | sort | uniq -c
echo "== SF10 parity of the changed paths vs the reference module"
timeout 300 python tools/run_tpch.py --sf 10 --check --reps 3 --trace --queries q13,q9,q16,q12,q3,q18 --out gpurun_out/sf10_check.json > gpurun_out/sf10_check.log 2>&1; echo "rc=$?"
grep -o '"query": "[a-z0-9]*"\|"parity": "[^"]*"\|"device_ms_min": [0-9.]*' gpurun_out/sf10_check.log | paste - - - | head -8
python tools/show_tpch.py gpurun_out/sf100_n1_all22.json profiles/r01_tpch_sf100_n1_all22_v3.json 2>/dev/null | tail -24
du -sh gpurun_out
