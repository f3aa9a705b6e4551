#!/bin/bash
# round 2, GPU-box visit 1: A/B of the opt-in code-generator switches written at the end of round 1 (never measured)
# against the default build on identical device-generated data; ncu of q1_k0 for default / tier0smem / next3
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt; nvidia-smi -L >> gpurun_out/nproc.txt
V="default,idx32,runagg,tier0smem,next3,mat,pack32"
ALL="q1,q2,q3,q4,q5,q6,q7,q8,q9,q10,q11,q12,q13,q14,q15,q16,q17,q18,q19,q20,q21,q22"
timeout 500 python tools/ab_variants.py --sf 10 --device-gen --reps 5 --variants $V --queries $ALL --out gpurun_out/r02_ab_candidates_sf10.json > gpurun_out/r02_ab_candidates_sf10.log 2>&1; echo "rc=$?"
timeout 500 python tools/ab_variants.py --sf 100 --device-gen --reps 3 --variants $V --queries q1,q6,q5,q9,q10,q18,q21,q14,q19,q3 --out gpurun_out/r02_ab_candidates_sf100.json > gpurun_out/r02_ab_candidates_sf100.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_ab_candidates_sf10", "gpurun_out/r02_ab_candidates_sf100"):
    try:
        r = json.load(open(f + ".json"))
    except Exception:
        r = [json.loads(l) for l in open(f + ".log") if l.startswith("{")]
    qs = []
    for x in r:
        if x["query"] not in qs: qs.append(x["query"])
    print(f)
    for q in qs:
        print(" ", q, "  ".join("%s %.3f (%s)" % (x["variant"], x["device_ms_min"], x["vs_first_variant"][:8]) for x in r if x["query"] == q))
PY
cap() {  # name regex skip query so
    SDQLB200_SO=$5 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 \
        -o gpurun_out/$1 -f python tools/run_tpch.py --sf 10 --device-gen --queries $4 --reps 1 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
    python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu.txt 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
    python tools/ncu_hot.py gpurun_out/$1_source.csv 40 > gpurun_out/$1_hot.txt 2>&1
    rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_source.csv
    head -22 gpurun_out/$1_ncu.txt
}
cap r02_q1_k0_tier0smem "^q1_k0" 1 q1 gpurun_variants/tier0smem.so
cap r02_q1_k0_next3 "^q1_k0" 1 q1 gpurun_variants/next3.so
du -sh gpurun_out
