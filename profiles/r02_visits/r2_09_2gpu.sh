#!/bin/bash
# round 2, visit 9 (2 GPUs): after the fixes of visit 6 (result-row gather with host columns, stale CUDA errors, counted +
# re-planned merged tables): exchange library tests under both bootstraps, SF100 bench line at N=2 (all 22 with parity),
# forced-hash Q1/Q3/Q5/Q9/Q18 at SF10
set -u
mkdir -p gpurun_out
( time timeout 1300 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_v9_tests_multi.log 2>&1; echo "multi rc=$?"; tail -12 gpurun_out/r02_v9_tests_multi.log | cut -c1-600
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_v9_bench_sf100_n2.json 2> gpurun_out/r02_v9_bench_sf100_n2.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_v9_bench_sf100_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'result_check', 'all_queries_ms')}); print(d['roofline']['frac'], d['e2e']); print(d['detail']['merge'])
for q, v in d['per_query'].items(): print(' ', q, v.get('ms'), v.get('parity', v.get('error'))[:90])"
tail -5 gpurun_out/r02_v9_bench_sf100_n2.err | cut -c1-300
SDQLB200_FORCE_HASH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/run_tpch_dist.py --sf 10 --device-gen --queries q1,q3,q5,q9,q18 --reps 3 --check ref --out gpurun_out/r02_v9_sf10_n2_forcehash.json > gpurun_out/r02_v9_sf10_n2_forcehash.log 2>&1; echo "forcehash rc=$?"; grep '^{' gpurun_out/r02_v9_sf10_n2_forcehash.log | cut -c1-420
