#!/bin/bash
# GPU-box visit 12: parity, then A/B of rows-per-thread and string-search variants on identical data
set -u
mkdir -p gpurun_out
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -4 gpurun_out/tests.log
echo "== A/B"
timeout 1500 python tools/ab_variants.py --sf 10 --reps 5 --queries q2,q3,q4,q5,q7,q8,q9,q10,q11,q12,q13,q14,q15,q16,q17,q18,q19,q20,q21,q22 --variants default,rows4,strbytes --out gpurun_out/ab12.json > gpurun_out/ab12.log 2>&1; echo rc=$?
python - <<'PY'
import json
r = json.load(open("gpurun_out/ab12.json"))
qs = []
for x in r:
    if x["query"] not in qs: qs.append(x["query"])
for q in qs:
    print(q, "  ".join("%s %.3f (%s)" % (x["variant"], x["device_ms_min"], x["vs_first_variant"][:12]) for x in r if x["query"] == q))
PY
