#!/bin/bash
# round 2, visit 17 (1 GPU): shared-memory Bloom summaries for selective single-key probes + the queue-skip vote: GPU suite,
# A/B against the build without the summaries on all 22 queries at SF100, ncu of q17_k1
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v17_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_v17_tests_gpu.log | cut -c1-400
echo "== bloom A/B SF100"
ALL="q1,q2,q3,q4,q5,q6,q7,q8,q9,q10,q11,q12,q13,q14,q15,q16,q17,q18,q19,q20,q21,q22"
SDQLB200_DEBUG=1 timeout 600 python tools/ab_variants.py --sf 100 --device-gen --reps 3 --variants default,nobloom --queries $ALL --out gpurun_out/r02_v17_ab_bloom_sf100.json > gpurun_out/r02_v17_ab_bloom_sf100.log 2> gpurun_out/r02_v17_ab_bloom_sf100.err; echo "rc=$?"
grep "Bloom summary" gpurun_out/r02_v17_ab_bloom_sf100.err | sort -u > gpurun_out/r02_v17_blooms.txt; cat gpurun_out/r02_v17_blooms.txt | cut -c1-120
grep '^{' gpurun_out/r02_v17_ab_bloom_sf100.log | python -c "
import sys, json
rows = [json.loads(l) for l in sys.stdin]
qs = []
for x in rows:
    if x['query'] not in qs: qs.append(x['query'])
tot = {}
for q in qs:
    r = {x['variant']: x for x in rows if x['query'] == q}
    for v, x in r.items(): tot[v] = tot.get(v, 0) + x['device_ms_min']
    print(' ', q, '  '.join('%s %.3f (%s)' % (v, x['device_ms_min'], x['vs_first_variant'][:4]) for v, x in r.items()), r['default']['kernels'])
print('  totals', tot)"
echo "== ncu q17_k1 SF10"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:^q17_k1" --launch-skip 2 --launch-count 1 -o gpurun_out/r02_q17_k1 -f python tools/run_tpch.py --sf 100 --device-gen --queries q17 --reps 2 > gpurun_out/ncu_r02_q17_k1.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r02_q17_k1.ncu-rep > gpurun_out/r02_v17_q17_k1_ncu.txt 2>&1
ncu -i gpurun_out/r02_q17_k1.ncu-rep --page source --csv > gpurun_out/r02_q17_k1_source.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/r02_q17_k1_source.csv 30 > gpurun_out/r02_v17_q17_k1_hot.txt 2>&1
rm -f gpurun_out/r02_q17_k1.ncu-rep gpurun_out/r02_q17_k1_source.csv
head -24 gpurun_out/r02_v17_q17_k1_ncu.txt
du -sh gpurun_out
