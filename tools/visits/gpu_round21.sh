#!/bin/bash
# GPU-box visit 21 (last of the round): all 22 queries at SF10 with the final default build, every result compared with
# the reference module (oracle/_ref, 16 host threads) on the same generated inputs
set -u
mkdir -p gpurun_out
Q="q12,q13,q9,q16,q5,q3,q7,q8,q10,q17,q19,q20,q18,q21,q4,q1,q6,q14,q15,q2,q11,q22"
timeout 280 python tools/run_tpch.py --sf 10 --check --reps 3 --queries $Q --out gpurun_out/sf10_all22_v5_checked.json > gpurun_out/sf10_all22_v5_checked.log 2>&1; echo "rc=$?"
grep -o '"query": "[a-z0-9]*"\|"device_ms_min": [0-9.]*\|"ref_ms": [0-9.]*\|"parity": "[^"]*"' gpurun_out/sf10_all22_v5_checked.log | paste - - - -
