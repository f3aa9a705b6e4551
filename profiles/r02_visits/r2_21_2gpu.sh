#!/bin/bash
# round 2, visit 21 (2 GPUs): the final build at N = 2 (default bench line, all 22 with parity) + the exchange tests
set -u
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_v21_bench_sf100_n2.json 2> gpurun_out/r02_v21_bench_sf100_n2.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_v21_bench_sf100_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'result_check', 'all_queries_ms')}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']); print(d['detail']['merge'])
print({q: (round(v.get('ms', -1), 3), (v.get('parity') or v.get('error'))[:2]) for q, v in d['per_query'].items()})"
tail -3 gpurun_out/r02_v21_bench_sf100_n2.err | cut -c1-300
( time timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_v21_tests_multi.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/r02_v21_tests_multi.log | cut -c1-300
nproc
