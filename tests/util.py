"""shared helpers for the parity tests."""
import json
import os

from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QUERIES = ["q%d" % i for i in range(1, 23)]
QUERY_SCRIPT = os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py")


def golden(sf):
    path = os.path.join(ROOT, "tests", "golden", "tpch_sf%s.json" % ("%g" % sf).replace(".", "p"))
    raw = json.load(open(path))
    out = {}
    for q, r in raw["queries"].items():
        if isinstance(r, list) and len(r) == 2 and r[0] == "f" and isinstance(r[1], str):
            out[q] = float(r[1])
        else:
            out[q] = [tuple(float(v[1]) if isinstance(v, list) else v for v in row) for row in r]
    return out


_cache = {}


def compact_db(sf, query_args):
    """db for the B200 backend: per relation the generator's compact columns in schema order (None = not generated)."""
    g = _cache.setdefault(sf, {"g": TPCH(sf), "t": {}})
    for t in query_args:
        if t not in g["t"]:
            cols = g["g"].columns(t)
            g["t"][t] = [cols.get(c) for c, _ in SCHEMAS[t]]
    return [g["t"][t] for t in query_args]


def ref_db(sf, query_args):
    """db in the reference layout (int64 / float64 / <U n numpy arrays)."""
    g = _cache.setdefault(("ref", sf), {"g": TPCH(sf), "t": {}})
    for t in query_args:
        if t not in g["t"]:
            g["t"][t] = g["g"].ref_table(t, [c for c, _ in SCHEMAS[t]])
    return [g["t"][t] for t in query_args]
