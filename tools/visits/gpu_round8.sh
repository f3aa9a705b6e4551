#!/bin/bash
# GPU-box visit 8: parity with hit compaction, then A/B of code-generation variants on identical data
set -u
mkdir -p gpurun_out
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests.log
echo "== tests, every table counted / filtered" ; SDQLB200_BITS_MIN_BYTES=0 SDQLB200_COUNT_MIN_BYTES=0 SDQLB200_COUNT_MIN_RATIO=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/tests_forced.log 2>&1; echo "tests rc=$?" ; tail -3 gpurun_out/tests_forced.log
echo "== A/B"
timeout 1500 python tools/ab_variants.py --sf 10 --reps 5 --queries q2,q3,q4,q5,q7,q8,q9,q10,q12,q13,q14,q16,q17,q18,q19,q20,q21,q22 --variants default,nocompact,nostage,strw --out gpurun_out/ab8.json > gpurun_out/ab8.log 2>&1; echo rc=$?
python - <<'PY'
import json
r = json.load(open("gpurun_out/ab8.json"))
qs = []
for x in r:
    if x["query"] not in qs: qs.append(x["query"])
for q in qs:
    print(q, "  ".join("%s %.3f (%s)" % (x["variant"], x["device_ms_min"], x["vs_first_variant"][:12]) for x in r if x["query"] == q))
PY
