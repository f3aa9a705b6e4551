#!/bin/bash
# GPU-box visit 15 (short): ncu --set full of the two kernels the step traces single out -- the main pass of q12_k0
# (15 ms at SF100 for 3 M insertions) and q13_k0 (14.5 ms) -- at SF10, with per-instruction hot spots
set -u
mkdir -p gpurun_out
cap() {  # name regex skip query
    timeout 400 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 \
        -o gpurun_out/$1 -f python tools/run_tpch.py --sf 10 --device-gen --queries $4 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
    python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu.txt 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
    python tools/ncu_hot.py gpurun_out/$1_source.csv 30 > gpurun_out/$1_hot.txt 2>&1
    rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_source.csv
    head -24 gpurun_out/$1_ncu.txt
}
cap q12_k0_main "^q12_k0" 3 q12
cap q13_k0_textscan "^q13_k0" 1 q13
cap q9_k5 "^q9_k5" 1 q9
du -sh gpurun_out
