#!/bin/bash
# round 2, visit 13 (1 GPU): GPU suite; all 22 at SF100 with the deeper text-scan pipeline and the sector probe as default;
# ncu of q13_k0 (SF10) and of the .tbl parse kernel; default bench line
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v13_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_v13_tests_gpu.log | cut -c1-400
echo "== all 22 SF100"
timeout 600 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/r02_v13_sf100_all22.json > gpurun_out/r02_v13_sf100_all22.log 2> gpurun_out/r02_v13_sf100_all22.err; echo "rc=$?"
grep '^{' gpurun_out/r02_v13_sf100_all22.log | python -c "
import sys, json
tot = 0
for l in sys.stdin:
    x = json.loads(l); tot += x['device_ms_min']; print(' ', x['query'], '%.3f ms' % x['device_ms_min'], x.get('kernels'))
print('  total %.2f ms' % tot)"
echo "== ncu q13_k0 SF10"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:^q13_k0" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_q13_k0 -f python tools/run_tpch.py --sf 10 --device-gen --queries q13 --reps 2 > gpurun_out/ncu_r02_q13_k0.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r02_q13_k0.ncu-rep > gpurun_out/r02_v13_q13_k0_ncu.txt 2>&1
ncu -i gpurun_out/r02_q13_k0.ncu-rep --page source --csv > gpurun_out/r02_q13_k0_source.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/r02_q13_k0_source.csv 40 > gpurun_out/r02_v13_q13_k0_hot.txt 2>&1
rm -f gpurun_out/r02_q13_k0.ncu-rep gpurun_out/r02_q13_k0_source.csv
head -24 gpurun_out/r02_v13_q13_k0_ncu.txt
echo "== ncu .tbl parse"
timeout 300 ncu --set full --clock-control none --kernel-name "regex:k_tbl_parse" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_tbl_parse -f python tools/bench_tbl.py --mb 1024 --reps 2 > gpurun_out/ncu_r02_tbl_parse.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r02_tbl_parse.ncu-rep > gpurun_out/r02_v13_tbl_parse_ncu.txt 2>&1; rm -f gpurun_out/r02_tbl_parse.ncu-rep
head -24 gpurun_out/r02_v13_tbl_parse_ncu.txt
timeout 300 python tools/bench_tbl.py --mb 2048 --out gpurun_out/r02_v13_tbl_reader.json 2>&1 | tail -1 | cut -c1-1200
echo "== bench default"
( time timeout 900 python bench.py ) > gpurun_out/r02_v13_bench_sf100_n1.json 2> gpurun_out/r02_v13_bench_sf100_n1.err; echo "rc=$?"; cut -c1-2500 gpurun_out/r02_v13_bench_sf100_n1.json; tail -4 gpurun_out/r02_v13_bench_sf100_n1.err
echo "== reference arm"
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_v13_bench_ref.json 2> gpurun_out/r02_v13_bench_ref.err; echo "rc=$?"; cut -c1-900 gpurun_out/r02_v13_bench_ref.json
du -sh gpurun_out
