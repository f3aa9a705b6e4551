"""order-independent fingerprints of query results: the size-independent parity check at scale factors where the whole
result of the reference cannot be kept in git (SF10 / SF100).  A fingerprint is made FROM THE REFERENCE's result on the
generated data (tools/make_fingerprints.py, committed under tests/golden/) and compared with the fingerprint of the
CUDA path's result: row count and integer / string checksums exact, fp64 column sums within 1e-9 relative (of the sum of
magnitudes) -- the same bar as tests/compare.py, as a checksum of checksums."""
import math
import zlib

RTOL = 1e-9
MASK = (1 << 64) - 1


def _rows(res):
    if res is None:
        return None
    if isinstance(res, (int, float)):
        return [(res,)]
    return list(res.tuples()) if hasattr(res, "tuples") else list(res)


def fingerprint(res):
    """-> {"rows": n, "cols": [{"kind": "f", "sum": s, "abs": a} | {"kind": "x", "sum": checksum}, ...]}"""
    rows = _rows(res)
    if rows is None:
        return {"rows": -1, "cols": []}
    n = len(rows)
    if n == 0:
        return {"rows": 0, "cols": []}
    cols = []
    for j in range(len(rows[0])):
        v0 = rows[0][j]
        if isinstance(v0, float):
            vals = [float(r[j]) for r in rows]
            cols.append({"kind": "f", "sum": math.fsum(vals), "abs": math.fsum(abs(v) for v in vals)})
        else:
            s = 0
            for r in rows:
                v = r[j]
                if isinstance(v, str):
                    v = zlib.crc32(v.rstrip("\x00").encode("latin1", "replace"))
                elif isinstance(v, bool):
                    v = int(v)
                s = (s + (int(v) * 0x9e3779b97f4a7c15 & MASK)) & MASK
            cols.append({"kind": "x", "sum": s})
    return {"rows": n, "cols": cols}


def match(mine, ref, rtol=RTOL):
    """-> None if the fingerprints agree, else a short description of the first difference"""
    if mine["rows"] != ref["rows"]:
        return "row count %d != %d" % (mine["rows"], ref["rows"])
    if len(mine["cols"]) != len(ref["cols"]):
        return "%d fields != %d" % (len(mine["cols"]), len(ref["cols"]))
    for j, (a, b) in enumerate(zip(mine["cols"], ref["cols"])):
        if a["kind"] != b["kind"]:
            return "field %d: kind %s != %s" % (j, a["kind"], b["kind"])
        if a["kind"] == "x":
            if a["sum"] != b["sum"]:
                return "field %d: integer / string checksum differs" % j
        else:
            scale = max(abs(a["abs"]), abs(b["abs"]), 1e-300)
            if abs(a["sum"] - b["sum"]) > rtol * scale or abs(a["abs"] - b["abs"]) > rtol * scale:
                return "field %d: fp64 sum %r != %r" % (j, a["sum"], b["sum"])
    return None
