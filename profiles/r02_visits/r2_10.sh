#!/bin/bash
# round 2, visit 10 (1 GPU): GPU suite after the text-path rewrite (chunk-row marks, per-row resolve), the packed-first-part
# presence filter and the >= 4 GB count rule; sector-wide hash probe A/B; all 22 at SF100; ncu q13_k0 / q9_k5; default bench
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v10_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_v10_tests_gpu.log | cut -c1-400
echo "== probe A/B SF100"
timeout 400 python tools/ab_variants.py --sf 100 --device-gen --reps 5 --variants default,probesector --queries q9,q16,q20,q2,q11 --out gpurun_out/r02_v10_ab_probe_sf100.json > gpurun_out/r02_v10_ab_probe_sf100.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r02_v10_ab_probe_sf100.log | python -c "
import sys, json
for l in sys.stdin:
    x = json.loads(l); print(' ', x['query'], x['variant'], '%.3f' % x['device_ms_min'], x['vs_first_variant'][:40], x['kernels'])"
echo "== all 22 SF100"
SDQLB200_DEBUG=1 timeout 600 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/r02_v10_sf100_all22.json > gpurun_out/r02_v10_sf100_all22.log 2> gpurun_out/r02_v10_sf100_all22.err; echo "rc=$?"
grep "rows reach\|key domain" gpurun_out/r02_v10_sf100_all22.err | sort -u > gpurun_out/r02_v10_sf100_replans.txt
grep '^{' gpurun_out/r02_v10_sf100_all22.log | python -c "
import sys, json
tot = 0
for l in sys.stdin:
    x = json.loads(l); tot += x['device_ms_min']; print(' ', x['query'], '%.3f ms' % x['device_ms_min'], x.get('kernels'))
print('  total %.2f ms' % tot)"
echo "== ncu q13_k0 / q9_k5 SF10"
for KQ in q13_k0:q13 q9_k5:q9; do
  K=${KQ%%:*}; Q=${KQ##*:}
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:^$K" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_$K -f python tools/run_tpch.py --sf 10 --device-gen --queries $Q --reps 2 > gpurun_out/ncu_r02_$K.log 2>&1; echo "ncu $K rc=$?"
  python tools/ncu_summary.py gpurun_out/r02_$K.ncu-rep > gpurun_out/r02_v10_${K}_ncu.txt 2>&1
  ncu -i gpurun_out/r02_$K.ncu-rep --page source --csv > gpurun_out/r02_${K}_source.csv 2>/dev/null
  python tools/ncu_hot.py gpurun_out/r02_${K}_source.csv 50 > gpurun_out/r02_v10_${K}_hot.txt 2>&1
  rm -f gpurun_out/r02_$K.ncu-rep gpurun_out/r02_${K}_source.csv
  head -24 gpurun_out/r02_v10_${K}_ncu.txt
done
echo "== bench default"
( time timeout 900 python bench.py ) > gpurun_out/r02_v10_bench_sf100_n1.json 2> gpurun_out/r02_v10_bench_sf100_n1.err; echo "rc=$?"; cut -c1-3000 gpurun_out/r02_v10_bench_sf100_n1.json; tail -6 gpurun_out/r02_v10_bench_sf100_n1.err
du -sh gpurun_out
