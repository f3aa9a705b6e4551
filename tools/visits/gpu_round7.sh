#!/bin/bash
# GPU-box visit 7: SF10 all-22 timings (fact tables generated on the device) with the table-side decisions logged,
# ncu captures of the join / string kernels condensed on the box (reports are too big to bring back)
set -u
mkdir -p gpurun_out
echo "== default"; SDQLB200_DEBUG=1 timeout 600 python tools/run_tpch.py --sf 10 --device-gen --out gpurun_out/ab_default.json > gpurun_out/ab_default.log 2> gpurun_out/ab_default.err; echo rc=$?
grep "sdqlb200" gpurun_out/ab_default.err | sort | uniq -c | sort -rn | head -40
echo "== no presence bits"; SDQLB200_BITS_MIN_BYTES=1000000000000 timeout 600 python tools/run_tpch.py --sf 10 --device-gen --out gpurun_out/ab_nobits.json > gpurun_out/ab_nobits.log 2>&1; echo rc=$?
python tools/show_tpch.py gpurun_out/ab_default.json gpurun_out/ab_nobits.json
echo "== ncu"
ncu_one() {  # name, query, kernel regex
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:$3" --launch-skip 1 --launch-count 1 -o gpurun_out/$1 -f python tools/run_tpch.py --sf 10 --device-gen --queries $2 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
  cat gpurun_out/$1_ncu.txt
}
ncu_one q13_k0 q13 "^q13_k0"
ncu_one q5_k5 q5 "^q5_k5"
ncu_one q18_k0 q18 "^q18_k0"
ncu_one q9_k5 q9 "^q9_k5"
ncu_one q3_k2 q3 "^q3_k2"
ncu_one q21_k5 q21 "^q21_k5"
du -sh gpurun_out
