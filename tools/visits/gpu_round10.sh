#!/bin/bash
# GPU-box visit 10: parity (normal + forced table paths), SF10 / SF100 timings with device-generated fact tables
set -u
mkdir -p gpurun_out
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -4 gpurun_out/tests.log
echo "== tests, every table counted / filtered" ; SDQLB200_BITS_MIN_BYTES=0 SDQLB200_COUNT_MIN_BYTES=0 SDQLB200_COUNT_MIN_RATIO=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/tests_forced.log 2>&1; echo "tests rc=$?" ; tail -3 gpurun_out/tests_forced.log
echo "== SF10"; SDQLB200_DEBUG=1 timeout 600 python tools/run_tpch.py --sf 10 --device-gen --out gpurun_out/sf10_dg.json > gpurun_out/sf10_dg.log 2> gpurun_out/sf10_dg.err; echo rc=$?
grep "sdqlb200" gpurun_out/sf10_dg.err | sort | uniq -c | sort -rn | head
echo "== SF100"; timeout 900 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/sf100_dg.json > gpurun_out/sf100_dg.log 2>&1; echo rc=$?
