#!/bin/bash
# GPU-box visit 4: parity tests + SF10 (checked against the reference module) + SF100 timings after a codegen change
set -u
mkdir -p gpurun_out
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed"; exit 1; fi
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests.log
echo "== tests, every table counted / filtered" ; SDQLB200_BITS_MIN_BYTES=0 SDQLB200_COUNT_MIN_BYTES=0 SDQLB200_COUNT_MIN_RATIO=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/tests_forced.log 2>&1; echo "tests rc=$?" ; tail -3 gpurun_out/tests_forced.log
echo "== SF10 all 22 (checked)"
timeout 1500 python tools/run_tpch.py --sf 10 ${CHECK:---check} --out gpurun_out/sf10_all22.json > gpurun_out/sf10_all22.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/sf10_all22.log | tail -24
echo "== SF100 all 22, one GPU"
timeout 1500 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/sf100_n1_all22.json > gpurun_out/sf100_n1_all22.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/sf100_n1_all22.log | tail -26
