#!/bin/bash
# round 2, visit 16 (1 GPU): host-side narrowing in front of the upload (SDQLB200_HOST_NARROW): GPU test of the path, then
# the end-to-end step with and without it (same process layout as the bench: plain numpy columns, store disabled)
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_ingest.py -m gpu -x -q ) > gpurun_out/r02_v16_tests_ingest.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_v16_tests_ingest.log | cut -c1-600
for HN in 0 1; do
  ( time SDQLB200_HOST_NARROW=$HN timeout 600 python bench.py --queries none --no-cpu-baseline --e2e-steps 4 ) > gpurun_out/r02_v16_bench_hn$HN.json 2> gpurun_out/r02_v16_bench_hn$HN.err; echo "bench HOST_NARROW=$HN rc=$?"
  grep '^{' gpurun_out/r02_v16_bench_hn$HN.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); e = d['e2e']; print({k: e.get(k) for k in ('value', 'ms_per_step', 'h2d_bytes_per_step', 'result', 'error')}, d['value'])"
  tail -3 gpurun_out/r02_v16_bench_hn$HN.err | cut -c1-300
done
nproc
