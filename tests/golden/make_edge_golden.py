#!/usr/bin/env python3
"""Golden vectors for edge-case inputs: outputs of the REAL reference (oracle/_ref, 1 thread) for all 22 queries on
  * empty relations (every table has zero rows)
  * ragged relations: the first 257 orders of SF0.01 and exactly their lineitems -- row counts that are no multiple of
    4 / 32 / 128 (the vector width, the warp, the rows one warp examines per iteration); dimension tables complete
Run in the dev container (needs /root/reference for the build): python tests/golden/make_edge_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import build_ref  # noqa: E402
import ref_runner as rr  # noqa: E402
from sdqlpy_b200.tpch.gen import SCHEMAS, SEED, TPCH  # noqa: E402
from util import ragged_rows  # noqa: E402


def dump(mod, tabs, tag, extra):
    out = dict(extra, seed=SEED, queries={})
    for q in rr.QUERY_ARGS:
        r = rr.run(mod, q, [tabs[t] for t in rr.QUERY_ARGS[q]])
        if isinstance(r, list):
            r = sorted(r, key=repr)
            r = [[("f", repr(v)) if isinstance(v, float) else v for v in row] for row in r]
        else:
            r = ("f", repr(r))
        out["queries"][q] = r
    path = os.path.join(ROOT, "tests", "golden", "tpch_%s.json" % tag)
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print(path, os.path.getsize(path))


def main():
    name = build_ref.build(1, 1)
    mod = rr.load(name)
    g = TPCH(0.01)
    full = {t: g.ref_table(t, [c for c, _ in SCHEMAS[t]]) for t in SCHEMAS}
    dump(mod, {t: [c[:0] for c in cols] for t, cols in full.items()}, "empty", {"sf": 0, "reference_module": name})
    rows = ragged_rows(full["orders"][0], full["lineitem"][0])
    assert rows["lineitem"] % 4 and rows["orders"] % 4
    dump(mod, {t: [c[:rows.get(t)] for c in cols] for t, cols in full.items()}, "ragged",
         {"sf": 0.01, "rows": rows, "reference_module": name})


if __name__ == "__main__":
    main()
