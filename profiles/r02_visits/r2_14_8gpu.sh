#!/bin/bash
# round 2, visit 14 (8 GPUs): the headline configuration -- SF100 strong-scaled over 8 B200s, all 22 queries with parity
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_v14_topo8.txt 2>&1; nproc > gpurun_out/r02_v14_host.txt; free -g >> gpurun_out/r02_v14_host.txt
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r02_v14_bench_sf100_n8.json 2> gpurun_out/r02_v14_bench_sf100_n8.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_v14_bench_sf100_n8.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'result_check', 'all_queries_ms')}); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step']); print(d['detail'])
for q, v in d['per_query'].items(): print(' ', q, v.get('ms'), (v.get('parity') or v.get('error'))[:60])"
tail -8 gpurun_out/r02_v14_bench_sf100_n8.err | cut -c1-300
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 20 --warmup 3 --queries q1,q6,q3,q5,q9,q18 ) > gpurun_out/r02_v14_bench_sf100_n4.json 2> gpurun_out/r02_v14_bench_sf100_n4.err; echo "bench4 rc=$?"; grep '^{' gpurun_out/r02_v14_bench_sf100_n4.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'result_check')}, d['e2e']['value'], d['per_query_ms'])"
