#!/bin/bash
# GPU-box visit (not run yet): A/B of the opt-in code-generator switches against the default build on identical
# device-generated data, results compared between builds; SF10 for all 22 queries, SF100 for the queries each switch targets
set -u
mkdir -p gpurun_out
V="default,idx32,runagg,tier0smem,next3,mat,pack32"
ALL="q1,q2,q3,q4,q5,q6,q7,q8,q9,q10,q11,q12,q13,q14,q15,q16,q17,q18,q19,q20,q21,q22"
timeout 600 python tools/ab_variants.py --sf 10 --device-gen --reps 5 --variants $V --queries $ALL --out gpurun_out/ab_candidates_sf10.json > gpurun_out/ab_candidates_sf10.log 2>&1; echo "rc=$?"
timeout 600 python tools/ab_variants.py --sf 100 --device-gen --reps 3 --variants $V --queries q1,q6,q5,q9,q10,q18,q21,q14,q19 --out gpurun_out/ab_candidates_sf100.json > gpurun_out/ab_candidates_sf100.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/ab_candidates_sf10", "gpurun_out/ab_candidates_sf100"):
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
