#!/bin/bash
# round 2, visit 8 (1 GPU): GPU suite with the run-time key domains + aligned text scan; cardinality-pass threshold A/B
# (SDQLB200_COUNT_MIN_RATIO 4 / 2 / 1, all 22 at SF100); ncu of q13_k0 and q9_k5; e2e probe after the ingest fix; the full
# default bench line; .tbl reader throughput
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_v8_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_v8_tests_gpu.log | cut -c1-400
echo "== count ratio A/B"
for R in 4 2 1; do
  SDQLB200_COUNT_MIN_RATIO=$R timeout 400 python tools/run_tpch.py --sf 100 --device-gen --reps 3 --out gpurun_out/r02_v8_sf100_ratio$R.json > gpurun_out/r02_v8_sf100_ratio$R.log 2>&1; echo "ratio $R rc=$?"
done
python - <<'PY'
import json
def load(r):
    try: return {x["query"]: x for x in json.load(open("gpurun_out/r02_v8_sf100_ratio%d.json" % r))}
    except Exception as e: print("ratio", r, e); return {}
d = {r: load(r) for r in (4, 2, 1)}
tot = {r: 0.0 for r in d}
for q in d[4]:
    print(" ", q, "  ".join("r%d %.3f" % (r, d[r][q]["device_ms_min"]) for r in d if q in d[r]), d[4][q].get("kernels"))
    for r in d:
        if q in d[r]: tot[r] += d[r][q]["device_ms_min"]
print("  totals", tot)
PY
echo "== ncu q13_k0 / q9_k5 SF10"
for KQ in q13_k0:q13 q9_k5:q9; do
  K=${KQ%%:*}; Q=${KQ##*:}
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:^$K" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_$K -f python tools/run_tpch.py --sf 10 --device-gen --queries $Q --reps 2 > gpurun_out/ncu_r02_$K.log 2>&1; echo "ncu $K rc=$?"
  python tools/ncu_summary.py gpurun_out/r02_$K.ncu-rep > gpurun_out/r02_${K}_ncu.txt 2>&1
  ncu -i gpurun_out/r02_$K.ncu-rep --page source --csv > gpurun_out/r02_${K}_source.csv 2>/dev/null
  python tools/ncu_hot.py gpurun_out/r02_${K}_source.csv 50 > gpurun_out/r02_${K}_hot.txt 2>&1
  rm -f gpurun_out/r02_$K.ncu-rep gpurun_out/r02_${K}_source.csv
  head -24 gpurun_out/r02_${K}_ncu.txt
done
echo "== e2e probe SF100"
timeout 400 python tools/e2e_probe.py --sf 100 --out gpurun_out/r02_v8_e2e_probe_sf100.json 2>&1 | grep '^{' | cut -c1-300
echo "== tbl reader"
timeout 300 python tools/bench_tbl.py --mb 2048 --out gpurun_out/r02_v8_tbl_reader.json 2>&1 | tail -3 | cut -c1-1200
echo "== bench default"
( time timeout 900 python bench.py ) > gpurun_out/r02_v8_bench_sf100_n1.json 2> gpurun_out/r02_v8_bench_sf100_n1.err; echo "rc=$?"; cut -c1-6000 gpurun_out/r02_v8_bench_sf100_n1.json; tail -6 gpurun_out/r02_v8_bench_sf100_n1.err
du -sh gpurun_out
