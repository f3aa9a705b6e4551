#!/bin/bash
# one GPU-box visit: parity tests, bench line, pipeline A/B at SF10, ncu launch list + full capture of the Q1 kernel
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping to save GPU minutes"; exit 1; fi
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cat gpurun_out/bench_q1.json
echo "== A/B variants"
timeout 900 python tools/ab_variants.py --sf 10 --queries q1,q6,q3,q5,q9,q18 --variants default,legacy,ring4,ring2 --out gpurun_out/ab.json > gpurun_out/ab.log 2>&1; echo "rc=$?"; cat gpurun_out/ab.log | cut -c1-400
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full q1_k0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q1_k0 -s 3 -c 1 -o gpurun_out/q1_k0 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
