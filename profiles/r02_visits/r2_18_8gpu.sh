#!/bin/bash
# round 2, visit 18 (8 GPUs): the headline configuration again with the final build -- ranks pinned to core slices,
# host-side narrowing in the end-to-end step; all 22 queries with parity
set -u
mkdir -p gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r02_v18_bench_sf100_n8.json 2> gpurun_out/r02_v18_bench_sf100_n8.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_v18_bench_sf100_n8.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'result_check', 'all_queries_ms')}); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline'].get('kernel_share_of_step'), d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']); print(d['detail'].get('dimension_tables_host_s'), d['clocks'])
for q, v in d['per_query'].items(): print(' ', q, v.get('ms'), (v.get('parity') or v.get('error'))[:40])"
tail -6 gpurun_out/r02_v18_bench_sf100_n8.err | cut -c1-300
