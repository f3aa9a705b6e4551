#!/usr/bin/env python3
"""Golden vectors for the device-side .tbl reader: a small lineitem / orders / customer .tbl fixture and the columns
the REAL reference read_csv (sdql_lib.py:118-128, imported from /root/reference -- only in the dev container) makes
of it.  Writes tests/golden/tbl_fixture.json: {"tables": {name: {"text": <.tbl text>, "columns": {col: values}}}}.
Floats are stored as hex (float.hex) so the comparison is bit exact.   python tests/golden/make_tbl_golden.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

from sdqlpy import sdql_lib as ref  # noqa: E402  (the reference)
from sdqlpy_b200 import tbl  # noqa: E402
from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH  # noqa: E402


def ref_schema(schema):
    d = {}
    for name, kind in schema:
        d[name] = ref.string(kind[1]) if isinstance(kind, tuple) else {"int": int, "float": float, "date": ref.date}[kind]
    return {ref.record(d): bool}


def main():
    g = TPCH(0.001)
    out = {"tables": {}}
    for table, rows in (("lineitem", 60), ("orders", 40), ("customer", 25)):
        schema = SCHEMAS[table]
        cols = g.ref_table(table, [c for c, _ in schema])
        cols = [np.asarray(c)[:rows] if len(c) >= rows else np.asarray(c) for c in cols]
        text = tbl.format_tbl(schema, cols).decode("latin1")
        # a few hand-made rows on top of the generated ones: negative money, many decimals, long strings (cut to n)
        if table == "customer":
            text += "9001|Customer#X|an address that is much longer than the forty characters the schema allows|7|17-123-456-7890|-999.99|BUILDING|c|\n"
            text += "9002|Customer#Y|addr|24|34-000-000-0000|0.1|MACHINERY|0.30000000000000004 is not 0.3|\n"
            text += "9003|Customer#Z|addr|0|10-000-000-0000|123456789012.125|FURNITURE|big|\n"
        with tempfile.NamedTemporaryFile("w", suffix=".tbl", delete=False, newline="\n") as f:
            f.write(text)
            path = f.name
        r = ref.read_csv(path, ref_schema(schema), table)
        os.unlink(path)
        heads, data = r.getContainer()["headers"], r.getContainer()["data"]
        colsj = {}
        for (name, kind), a in zip(schema, data):
            if isinstance(kind, tuple):
                colsj[name] = [str(x) for x in a]
            elif kind == "float":
                colsj[name] = [float(x).hex() for x in a]
            else:
                colsj[name] = [int(x) for x in a]
        out["tables"][table] = {"text": text, "columns": colsj}
    p = os.path.join(ROOT, "tests", "golden", "tbl_fixture.json")
    json.dump(out, open(p, "w"))
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
