#!/bin/bash
# round 2, visit 20 (1 GPU): the final build as the driver will see it -- GPU suite, smoke(), default bench line, reference arm
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v20_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_v20_tests_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_v20_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_v20_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02_v20_bench_sf100_n1.json 2> gpurun_out/r02_v20_bench_sf100_n1.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_v20_bench_sf100_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'result_check', 'all_queries_ms')}); print(d['roofline']); print(d['roofline_q6']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']); print(d['cpu_baseline']); print(d['clocks'])
print({q: (round(v.get('ms', -1), 3), (v.get('parity') or v.get('error'))[:2]) for q, v in d['per_query'].items()})"
tail -3 gpurun_out/r02_v20_bench_sf100_n1.err | cut -c1-300
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_v20_bench_ref.json 2> gpurun_out/r02_v20_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_v20_bench_ref.json
