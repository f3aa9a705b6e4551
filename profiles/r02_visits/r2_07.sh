#!/bin/bash
# round 2, visit 7 (1 GPU): GPU test suite with the warp-cooperative text search + first-part presence filter; text A/B
# (default = resolve, noresolve, textaligned) at SF100; all 22 at SF100 with the table plans; where the plain-numpy e2e step
# spends its time; reference fingerprints at SF10 (all 22) and SF100 (the rest); ncu of q13_k0
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v7_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_v7_tests_gpu.log | cut -c1-400
echo "== text A/B SF100"
timeout 400 python tools/ab_variants.py --sf 100 --device-gen --reps 5 --variants default,noresolve,textaligned --queries q13,q9,q16 --out gpurun_out/r02_v7_ab_text_sf100.json > gpurun_out/r02_v7_ab_text_sf100.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r02_v7_ab_text_sf100.log | python -c "
import sys, json
for l in sys.stdin:
    x = json.loads(l); print(' ', x['query'], x['variant'], '%.3f' % x['device_ms_min'], x['vs_first_variant'][:40], x['kernels'])"
echo "== all 22 SF100"
SDQLB200_DEBUG=1 timeout 600 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/r02_v7_sf100_all22.json > gpurun_out/r02_v7_sf100_all22.log 2> gpurun_out/r02_v7_sf100_all22.err; echo "rc=$?"
grep "table\|rows reach" gpurun_out/r02_v7_sf100_all22.err | sort -u > gpurun_out/r02_v7_sf100_table_plans.txt
grep '^{' gpurun_out/r02_v7_sf100_all22.log | python -c "
import sys, json
tot = 0
for l in sys.stdin:
    x = json.loads(l); tot += x['device_ms_min']; print(' ', x['query'], '%.3f ms' % x['device_ms_min'], 'ws %.0f MB' % x['workspace_MB'])
print('  total %.2f ms' % tot)"
echo "== e2e probe"
timeout 300 python tools/e2e_probe.py --sf 10 --out gpurun_out/r02_v7_e2e_probe_sf10.json 2>&1 | grep '^{' | cut -c1-400
timeout 400 python tools/e2e_probe.py --sf 100 --out gpurun_out/r02_v7_e2e_probe_sf100.json 2>&1 | grep '^{' | cut -c1-400
echo "== ncu q13_k0 SF10"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:^q13_k0" --launch-skip 2 --launch-count 1 -o gpurun_out/r02_q13_k0 -f python tools/run_tpch.py --sf 10 --device-gen --queries q13 --reps 1 > gpurun_out/ncu_r02_q13_k0.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_q13_k0.ncu-rep > gpurun_out/r02_q13_k0_ncu.txt 2>&1
ncu -i gpurun_out/r02_q13_k0.ncu-rep --page source --csv > gpurun_out/r02_q13_k0_source.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/r02_q13_k0_source.csv 40 > gpurun_out/r02_q13_k0_hot.txt 2>&1
rm -f gpurun_out/r02_q13_k0.ncu-rep gpurun_out/r02_q13_k0_source.csv
head -24 gpurun_out/r02_q13_k0_ncu.txt
echo "== fingerprints SF10"; timeout 900 python tools/make_fingerprints.py --sf 10 --out gpurun_out/tpch_sf10_fingerprints.json --report gpurun_out/r02_v7_parity_sf10.json > gpurun_out/r02_v7_parity_sf10.log 2>&1; echo "rc=$?"; cut -c1-160 gpurun_out/r02_v7_parity_sf10.log
echo "== fingerprints SF100"; timeout 1500 python tools/make_fingerprints.py --sf 100 --queries q2,q4,q7,q8,q10,q11,q12,q13,q14,q15,q16,q17,q19,q20,q21,q22 --out gpurun_out/tpch_sf100_fingerprints_b.json --report gpurun_out/r02_v7_parity_sf100.json > gpurun_out/r02_v7_parity_sf100.log 2>&1; echo "rc=$?"; cut -c1-160 gpurun_out/r02_v7_parity_sf100.log
du -sh gpurun_out
