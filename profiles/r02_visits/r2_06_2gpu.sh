#!/bin/bash
# round 2, visit 6 (2 GPUs): the exchange library on hardware -- engine (one process, 2 GPUs) and process-per-GPU
# bootstraps, all 22 queries vs goldens, forced-hash runs (hash all-to-all on NVLink); SF100 bench line at N=2 (all 22 queries,
# strong scaling); Q9/Q18 at SF10 with every table hashed (table_merges > 0 at scale)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_tests_multi.log 2>&1; echo "multi rc=$?"; tail -30 gpurun_out/r02_tests_multi.log | cut -c1-600
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_bench_sf100_n2.json 2> gpurun_out/r02_bench_sf100_n2.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r02_bench_sf100_n2.json; tail -12 gpurun_out/r02_bench_sf100_n2.err
SDQLB200_FORCE_HASH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/run_tpch_dist.py --sf 10 --device-gen --queries q1,q3,q5,q9,q18 --reps 3 --out gpurun_out/r02_sf10_n2_forcehash.json > gpurun_out/r02_sf10_n2_forcehash.log 2>&1; echo "forcehash rc=$?"; tail -8 gpurun_out/r02_sf10_n2_forcehash.log | cut -c1-400
