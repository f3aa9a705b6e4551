#!/bin/bash
# round 2, visit 19 (4 GPUs): end-to-end step with / without the host-side narrowing at N = 4 (8 host threads per rank)
set -u
mkdir -p gpurun_out
for HN in auto 0; do
  ( time SDQLB200_HOST_NARROW=$HN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 --queries none --no-cpu-baseline --e2e-steps 4 ) > gpurun_out/r02_v19_bench_n4_hn$HN.json 2> gpurun_out/r02_v19_bench_n4_hn$HN.err; echo "bench HOST_NARROW=$HN rc=$?"
  grep '^{' gpurun_out/r02_v19_bench_n4_hn$HN.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); e = d['e2e']; print({k: e.get(k) for k in ('value', 'ms_per_step', 'h2d_bytes_per_step', 'result', 'error')}, d['value'], d['ms_per_step'])"
  tail -2 gpurun_out/r02_v19_bench_n4_hn$HN.err | cut -c1-200
done
nproc
