#!/bin/bash
# GPU-box visit 3: smoke, GPU parity tests, both bench arms, SF10 all-22 per-kernel times, SF100 all-22 on ONE GPU
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed"; exit 1; fi
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_q1.json
echo "== bench reference arm" ; timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_q1_ref.json 2> gpurun_out/bench_q1_ref.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_q1_ref.json
echo "== SF10 all 22"
timeout 900 python tools/run_tpch.py --sf 10 --out gpurun_out/sf10_all22.json > gpurun_out/sf10_all22.log 2>&1; echo "rc=$?"; cut -c1-260 gpurun_out/sf10_all22.log | tail -24
echo "== SF100 all 22, one GPU, fact tables generated in HBM"
timeout 1500 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/sf100_n1_all22.json > gpurun_out/sf100_n1_all22.log 2>&1; echo "rc=$?"; cut -c1-260 gpurun_out/sf100_n1_all22.log | tail -26
ls -la gpurun_out | head -40
