#!/bin/bash
# round 2, visit 23 (1 GPU, short): GPU parity tests (goldens, edge cases, C ABI) on the very last build
set -u
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_cabi.py tests/test_ingest.py -m gpu -x -q -k "not sf1 and not sf10" ) > gpurun_out/r02_v23_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_v23_tests.log | cut -c1-300
