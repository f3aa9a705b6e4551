#!/bin/bash
# round 2, visit 22 (1 GPU, short): unused scan columns no longer loaded (q21_k3/k4, q9_k4, q8_k6, q10_k1): parity + times
set -u
mkdir -p gpurun_out
( time timeout 200 python bench.py --queries q21,q9,q8,q10,q1,q3 --no-e2e --no-cpu-baseline --steps 5 ) > gpurun_out/r02_v22_bench.json 2> gpurun_out/r02_v22_bench.err; echo "rc=$?"; grep '^{' gpurun_out/r02_v22_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['roofline']['frac'], {q: (round(v.get('ms', -1), 3), (v.get('parity') or v.get('error'))[:2]) for q, v in d['per_query'].items()})"
tail -2 gpurun_out/r02_v22_bench.err | cut -c1-200
