#!/bin/bash
# GPU-box visit 13 (validation of the round's final state): parity tests (normal + forced table paths), both bench arms,
# SF10 all 22 checked against the reference module, SF100 all 22 on one GPU, ncu launch lists + full capture of q1_k0
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -4 gpurun_out/tests.log
echo "== tests, every table counted / filtered" ; SDQLB200_BITS_MIN_BYTES=0 SDQLB200_COUNT_MIN_BYTES=0 SDQLB200_COUNT_MIN_RATIO=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/tests_forced.log 2>&1; echo "tests rc=$?" ; tail -3 gpurun_out/tests_forced.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/bench_q1.json; tail -3 gpurun_out/bench_q1.err
echo "== bench reference arm" ; timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_q1_ref.json 2> gpurun_out/bench_q1_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_q1_ref.json
echo "== bench q6" ; timeout 600 python bench.py --query q6 --no-cpu-baseline > gpurun_out/bench_q6.json 2> gpurun_out/bench_q6.err; echo "rc=$?"; cut -c1-1200 gpurun_out/bench_q6.json
echo "== SF10 all 22 (checked)"
timeout 1500 python tools/run_tpch.py --sf 10 --check --out gpurun_out/sf10_all22.json > gpurun_out/sf10_all22.log 2>&1; echo "rc=$?"
echo "== SF100 all 22, one GPU"
timeout 1500 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/sf100_n1_all22.json > gpurun_out/sf100_n1_all22.log 2>&1; echo "rc=$?"
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/joins_launches.csv python tools/run_tpch.py --sf 10 --device-gen --queries q12,q9,q18,q13,q21 --reps 1 > gpurun_out/ncu_joins.log 2>&1; echo "rc=$?"
echo "== ncu full q1_k0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q1_k0 -s 3 -c 1 -o gpurun_out/q1_k0 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/q1_k0.ncu-rep > gpurun_out/q1_k0_ncu.txt 2>&1
rm -f gpurun_out/q1_k0.ncu-rep
du -sh gpurun_out
