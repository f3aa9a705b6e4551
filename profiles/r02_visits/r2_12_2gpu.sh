#!/bin/bash
# round 2, visit 12 (2 GPUs): exchange tests incl. the large-direct-table merges (sparse exchange / dense), per-step traces of
# the worst scalers after the cost-based merged count + sparse direct merge, SF100 bench line at N=2
set -u
mkdir -p gpurun_out
( time timeout 1300 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_v12_tests_multi.log 2>&1; echo "multi rc=$?"; tail -8 gpurun_out/r02_v12_tests_multi.log | cut -c1-600
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tools/run_tpch_dist.py --sf 100 --device-gen --queries q13,q17,q10,q21,q9,q20,q15,q3,q18 --reps 3 --trace --out gpurun_out/r02_v12_sf100_n2_trace.json > gpurun_out/r02_v12_sf100_n2_trace.log 2>&1; echo "trace rc=$?"; grep '^{' gpurun_out/r02_v12_sf100_n2_trace.log | python -c "
import sys, json
for l in sys.stdin:
    x = json.loads(l); print(' ', x['query'], 'device %.3f wall %.3f' % (x.get('device_ms', -1), x.get('latency_ms_wall', -1)), x.get('error', ''), x.get('trace_ms'))"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_v12_bench_sf100_n2.json 2> gpurun_out/r02_v12_bench_sf100_n2.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_v12_bench_sf100_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'result_check', 'all_queries_ms')}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']); print(d['detail']['merge'])
for q, v in d['per_query'].items(): print(' ', q, v.get('ms'), (v.get('parity') or v.get('error'))[:60])"
tail -5 gpurun_out/r02_v12_bench_sf100_n2.err | cut -c1-300
