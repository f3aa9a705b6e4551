#!/bin/bash
# round 2, visit 4 (1 GPU): full GPU test suite with the new defaults (incl. SF1 all-22 / SF10 headline parity against the
# live reference module and the device-side ingest), A/B of the scan pipelines for group-by kernels on all 22 queries
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02_tests_gpu.log
V="default,allreg,autol2"
ALL="q1,q2,q3,q4,q5,q6,q7,q8,q9,q10,q11,q12,q13,q14,q15,q16,q17,q18,q19,q20,q21,q22"
timeout 400 python tools/ab_variants.py --sf 10 --device-gen --reps 5 --variants $V --queries $ALL --out gpurun_out/r02_ab_pipes_sf10.json > gpurun_out/r02_ab_pipes_sf10.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
r = [json.loads(l) for l in open("gpurun_out/r02_ab_pipes_sf10.log") if l.startswith("{")]
qs = []
for x in r:
    if x["query"] not in qs: qs.append(x["query"])
for q in qs:
    print(" ", q, "  ".join("%s %.3f (%s)" % (x["variant"], x["device_ms_min"], x["vs_first_variant"][:8]) for x in r if x["query"] == q))
PY
