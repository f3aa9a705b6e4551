#!/bin/bash
# round 2, visit 2: q1_k0 pipelines -- register double-buffering (+ L2 prefetch) on top of the shared-memory tier-0
# accumulators and 32-bit indices, against the default and next3 builds, SF10 and SF100, ncu of the two new ones
set -u
mkdir -p gpurun_out
V="default,next3,q1a,q1d,q1e,q1f"
timeout 300 python tools/ab_variants.py --sf 10 --device-gen --reps 9 --variants $V --queries q1,q6 --out gpurun_out/r02_q1_pipes_sf10.json > gpurun_out/r02_q1_pipes_sf10.log 2>&1; echo "rc=$?"
timeout 300 python tools/ab_variants.py --sf 100 --device-gen --reps 5 --variants $V --queries q1,q6 --out gpurun_out/r02_q1_pipes_sf100.json > gpurun_out/r02_q1_pipes_sf100.log 2>&1; echo "rc=$?"
grep -h -o '"query": "[a-z0-9]*", "variant": "[a-z0-9]*", "sf": [0-9.]*, "device_ms_min": [0-9.]*' gpurun_out/r02_q1_pipes_sf10.log gpurun_out/r02_q1_pipes_sf100.log
cap() {  # name regex skip query so
    SDQLB200_SO=$5 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 \
        -o gpurun_out/$1 -f python tools/run_tpch.py --sf 10 --device-gen --queries $4 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
    python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu.txt 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
    python tools/ncu_hot.py gpurun_out/$1_source.csv 40 > gpurun_out/$1_hot.txt 2>&1
    rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_source.csv
    head -22 gpurun_out/$1_ncu.txt
}
cap r02_q1_k0_rega "^q1_k0" 1 q1 gpurun_variants/q1a.so
cap r02_q1_k0_regd "^q1_k0" 1 q1 gpurun_variants/q1d.so
