"""shared helpers for the parity tests."""
import json
import os

from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QUERIES = ["q%d" % i for i in range(1, 23)]
QUERY_SCRIPT = os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py")


# "ragged" edge-case inputs (tests/golden/make_edge_golden.py): the first RAGGED_ORDERS orders of SF0.01 and exactly their
# lineitems (row counts that are no multiple of 4 / 32 / 128), dimension tables complete -- referential integrity is
# kept because the reference aborts on a missing key (uncaught phmap at())
RAGGED_ORDERS = 257


def ragged_rows(orders_first_col, lineitem_first_col):
    """{table: rows to keep}: given the o_orderkey and l_orderkey columns (numpy arrays) of the full SF0.01 tables"""
    last = orders_first_col[RAGGED_ORDERS - 1]
    return {"orders": RAGGED_ORDERS, "lineitem": int((lineitem_first_col <= last).sum())}


def golden(sf):
    """sf: a scale factor (tpch_sf0p01.json ..) or the tag of an edge-case file ("empty", "ragged")"""
    tag = sf if isinstance(sf, str) else "sf%s" % ("%g" % sf).replace(".", "p")
    path = os.path.join(ROOT, "tests", "golden", "tpch_%s.json" % tag)
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


def cut_db(db, query_args, rows):
    """the relations of ``db`` (compact or reference layout) cut to ``rows[table]`` rows (0 = empty relations)"""
    from sdqlpy_b200.tpch.gen import Column
    out = []
    for t, rel in zip(query_args, db):
        n = rows.get(t) if isinstance(rows, dict) else rows   # None: the whole relation
        cut = []
        for c in rel:
            if c is None:
                cut.append(None)
            elif isinstance(c, Column):
                cut.append(Column(c.name, c.kind, c.data[:n], c.dictionary, c.width))
            else:
                cut.append(c[:n])
        out.append(cut)
    return out
