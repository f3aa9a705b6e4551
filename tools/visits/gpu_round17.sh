#!/bin/bash
# GPU-box visit 17 (validation of the current default build): full GPU test suite; A/B of the warp text scan (compact
# rare path) at SF10 and on Q13 at SF100; both bench arms; ncu launch list of the bench command
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== tests" ; timeout 480 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -3 gpurun_out/tests.log
echo "== A/B text scan, SF10"
timeout 200 python tools/ab_variants.py --sf 10 --reps 5 --queries q13,q9,q16 --variants default,notextscan --out gpurun_out/ab17.json > gpurun_out/ab17.log 2>&1; echo "rc=$?"
grep -o '"query": "[a-z0-9]*", "variant": "[a-z]*", "sf": 10.0, "device_ms_min": [0-9.]*' gpurun_out/ab17.log
grep -o '"vs_first_variant": "[^"]*"' gpurun_out/ab17.log | sort | uniq -c
echo "== Q13 SF100: default, then without the text scan"
timeout 150 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --queries q13,q9 --out gpurun_out/sf100_q13_default.json > gpurun_out/sf100_q13_default.log 2>&1; echo "rc=$?"
SDQLB200_SO=gpurun_variants/notextscan.so timeout 150 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --queries q13,q9 --out gpurun_out/sf100_q13_notextscan.json > gpurun_out/sf100_q13_notextscan.log 2>&1; echo "rc=$?"
grep -o '"query": "[a-z0-9]*"\|"device_ms_min": [0-9.]*\|"kernels": {[^}]*}' gpurun_out/sf100_q13_default.log gpurun_out/sf100_q13_notextscan.log | paste - - -
echo "== bench" ; timeout 400 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cut -c1-2200 gpurun_out/bench_q1.json; tail -2 gpurun_out/bench_q1.err
echo "== bench reference arm" ; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_q1_ref.json 2> gpurun_out/bench_q1_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_q1_ref.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
du -sh gpurun_out
