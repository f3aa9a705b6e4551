#!/usr/bin/env python3
"""Generates tests/golden/tpch_sf*.json: outputs of the REAL reference (oracle/_ref, built from /root/reference by
oracle/build_ref.py) for all 22 queries on the repo's synthetic TPC-H data.  Run in the dev container (needs
/root/reference for the build); the JSON travels with the repo so GPU-box tests never need the reference tree.

    python tests/golden/make_golden.py 0.01 0.05
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import build_ref  # noqa: E402
import ref_runner as rr  # noqa: E402
from sdqlpy_b200.tpch.gen import SCHEMAS, SEED, TPCH  # noqa: E402


def main():
    name = build_ref.build(1, 1)
    mod = rr.load(name)
    for sf in [float(x) for x in sys.argv[1:]] or [0.01]:
        g = TPCH(sf)
        tabs = {t: g.ref_table(t, [c for c, _ in SCHEMAS[t]]) for t in SCHEMAS}
        out = {"sf": sf, "seed": SEED, "reference_module": name, "queries": {}}
        for q in rr.QUERY_ARGS:
            r = rr.run(mod, q, [tabs[t] for t in rr.QUERY_ARGS[q]])
            if isinstance(r, list):
                r = sorted(r, key=repr)
                r = [[("f", repr(v)) if isinstance(v, float) else v for v in row] for row in r]
            else:
                r = ("f", repr(r))
            out["queries"][q] = r
        path = os.path.join(ROOT, "tests", "golden", "tpch_sf%s.json" % ("%g" % sf).replace(".", "p"))
        json.dump(out, open(path, "w"), separators=(",", ":"))
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
