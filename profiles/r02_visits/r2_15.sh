#!/bin/bash
# round 2, visit 15 (1 GPU): GPU suite; hash probe A/B (aligned pair = default vs whole sector); all 22 at SF100 with the
# counting build next to it (bytes-moved roofline); ncu launch list of the bench command; default bench + reference arm
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v15_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_v15_tests_gpu.log | cut -c1-400
echo "== probe A/B SF100"
timeout 400 python tools/ab_variants.py --sf 100 --device-gen --reps 5 --variants default,probe4 --queries q9,q20,q19,q17,q2,q16 --out gpurun_out/r02_v15_ab_probe_sf100.json > gpurun_out/r02_v15_ab_probe_sf100.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r02_v15_ab_probe_sf100.log | python -c "
import sys, json
for l in sys.stdin:
    x = json.loads(l); print(' ', x['query'], x['variant'], '%.3f' % x['device_ms_min'], x['vs_first_variant'][:40], x['kernels'])"
echo "== all 22 SF100 + counting build"
timeout 900 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --stats-so gpurun_variants/stats.so --out gpurun_out/r02_v15_sf100_all22.json > gpurun_out/r02_v15_sf100_all22.log 2> gpurun_out/r02_v15_sf100_all22.err; echo "rc=$?"
python tools/roofline_table.py gpurun_out/r02_v15_sf100_all22.json > gpurun_out/r02_v15_sf100_roofline_table.md 2>&1; cat gpurun_out/r02_v15_sf100_roofline_table.md
echo "== bench default"
( time timeout 900 python bench.py ) > gpurun_out/r02_v15_bench_sf100_n1.json 2> gpurun_out/r02_v15_bench_sf100_n1.err; echo "rc=$?"; cut -c1-2200 gpurun_out/r02_v15_bench_sf100_n1.json; tail -4 gpurun_out/r02_v15_bench_sf100_n1.err
echo "== reference arm"
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_v15_bench_ref.json 2> gpurun_out/r02_v15_bench_ref.err; echo "rc=$?"; cut -c1-500 gpurun_out/r02_v15_bench_ref.json
echo "== ncu launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v15_bench_launches.csv python bench.py --steps 4 --warmup 3 --queries none --no-e2e > gpurun_out/r02_v15_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
grep -c "q1_k0" gpurun_out/r02_v15_bench_launches.csv
du -sh gpurun_out
