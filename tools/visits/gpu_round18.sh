#!/bin/bash
# GPU-box visit 18 (short): ncu --set full of three narrow lineitem scans whose time does not follow their bytes
# (SF100: ~2.3 ms whether the kernel streams 4 or 28 bytes per row): q5_k5 (filter phase streams l_orderkey),
# q19_k1 (two 1-byte code columns), q17_k1
set -u
mkdir -p gpurun_out
cap() {  # name regex skip query
    timeout 300 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 \
        -o gpurun_out/$1 -f python tools/run_tpch.py --sf 10 --device-gen --queries $4 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
    python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu.txt 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
    python tools/ncu_hot.py gpurun_out/$1_source.csv 40 > gpurun_out/$1_hot.txt 2>&1
    rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_source.csv
    head -22 gpurun_out/$1_ncu.txt
}
cap q5_k5 "^q5_k5" 1 q5
cap q19_k1 "^q19_k1" 1 q19
cap q17_k1 "^q17_k1" 1 q17
du -sh gpurun_out
