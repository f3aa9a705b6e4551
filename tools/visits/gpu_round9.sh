#!/bin/bash
# GPU-box visit 9: parity tests (normal + every table path forced), bench (both arms), SF10 all 22 checked against the
# reference module, SF100 all 22 on one GPU, ncu of the compacted Q5 probe kernel
set -u
mkdir -p gpurun_out
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -4 gpurun_out/tests.log
echo "== tests, every table counted / filtered" ; SDQLB200_BITS_MIN_BYTES=0 SDQLB200_COUNT_MIN_BYTES=0 SDQLB200_COUNT_MIN_RATIO=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/tests_forced.log 2>&1; echo "tests rc=$?" ; tail -3 gpurun_out/tests_forced.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_q1.json; tail -3 gpurun_out/bench_q1.err
echo "== SF10 all 22 (checked)"
timeout 1500 python tools/run_tpch.py --sf 10 --check --out gpurun_out/sf10_all22.json > gpurun_out/sf10_all22.log 2>&1; echo "rc=$?"
echo "== SF100 all 22, one GPU"
timeout 1500 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/sf100_n1_all22.json > gpurun_out/sf100_n1_all22.log 2>&1; echo "rc=$?"
echo "== ncu q5_k5"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:^q5_k5" --launch-skip 1 --launch-count 1 -o gpurun_out/q5_k5 -f python tools/run_tpch.py --sf 10 --device-gen --queries q5 --reps 1 > gpurun_out/ncu_q5_k5.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/q5_k5.ncu-rep > gpurun_out/q5_k5_ncu.txt 2>&1
ncu -i gpurun_out/q5_k5.ncu-rep --page source --csv > gpurun_out/q5_k5_source.csv 2>/dev/null
rm -f gpurun_out/q5_k5.ncu-rep
