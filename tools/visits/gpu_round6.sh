#!/bin/bash
# GPU-box visit 6: A/B of the table-side options at SF10 (runtime knobs), ncu captures of the join kernels
set -u
mkdir -p gpurun_out
QS=q3,q5,q7,q8,q9,q10,q12,q17,q18,q20,q21
echo "== default"; timeout 600 python tools/run_tpch.py --sf 10 --device-gen --queries $QS --out gpurun_out/ab_default.json > gpurun_out/ab_default.log 2>&1; echo rc=$?
echo "== no presence bits"; SDQLB200_BITS_MIN_BYTES=1000000000000 timeout 600 python tools/run_tpch.py --sf 10 --device-gen --queries $QS --out gpurun_out/ab_nobits.json > gpurun_out/ab_nobits.log 2>&1; echo rc=$?
echo "== no bits, no cardinality pass"; SDQLB200_BITS_MIN_BYTES=1000000000000 SDQLB200_COUNT_MIN_BYTES=1000000000000 timeout 600 python tools/run_tpch.py --sf 10 --device-gen --queries $QS --out gpurun_out/ab_nobits_nocount.json > gpurun_out/ab_nobits_nocount.log 2>&1; echo rc=$?
python tools/show_tpch.py gpurun_out/ab_default.json gpurun_out/ab_nobits_nocount.json
python tools/show_tpch.py gpurun_out/ab_nobits.json gpurun_out/ab_nobits_nocount.json
echo "== ncu"
ncu_one() {  # name, query, kernel regex, extra env
  env $4 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:$3" --launch-skip 1 --launch-count 1 -o gpurun_out/$1 -f python tools/run_tpch.py --sf 10 --device-gen --queries $2 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
}
ncu_one q5_k5_bits q5 "^q5_k5" "X=1"
ncu_one q5_k5_nobits q5 "^q5_k5" "SDQLB200_BITS_MIN_BYTES=1000000000000"
ncu_one q3_k1_bits q3 "^q3_k1" "X=1"
ncu_one q18_k0 q18 "^q18_k0" "X=1"
ncu_one q9_k5 q9 "^q9_k5" "X=1"
ncu_one q13_k0 q13 "^q13_k0" "X=1"
ls -la gpurun_out/*.ncu-rep
