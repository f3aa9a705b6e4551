#!/bin/bash
# round 2, visit 3 (2 GPUs): the exchange library on hardware -- engine (one process, 2 GPUs) and process-per-GPU
# bootstraps, all 22 queries vs goldens, forced-hash runs (hash all-to-all on NVLink), old bench line at N=2
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_tests_multi.log 2>&1; echo "multi rc=$?"; tail -30 gpurun_out/r02_tests_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
