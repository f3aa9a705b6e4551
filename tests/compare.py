"""result comparison used by all parity tests: results are unordered sets of records; keys, counts and integer
fields must be bit-exact, fp64 fields agree within 1e-9 relative (BASELINE.json north_star)."""
import math

RTOL = 1e-9


def _split(row):
    exact = tuple(v for v in row if not isinstance(v, float))
    floats = tuple(v for v in row if isinstance(v, float))
    return exact, floats


def close(a, b, rtol=RTOL):
    return a == b or abs(a - b) <= rtol * max(abs(a), abs(b)) or (math.isnan(a) and math.isnan(b))


def compare(mine, ref, rtol=RTOL):
    """-> None if equal, else a short description of the first difference."""
    if isinstance(ref, (int, float)) or ref is None:
        if isinstance(mine, (int, float)) and ref is not None and close(float(mine), float(ref), rtol):
            return None
        return "scalar %r != %r" % (mine, ref)
    mine = list(mine.tuples()) if hasattr(mine, "tuples") else list(mine)
    ref = list(ref)
    if len(mine) != len(ref):
        return "row count %d != %d" % (len(mine), len(ref))
    idx = {}
    for r in ref:
        e, f = _split(r)
        idx.setdefault(e, []).append(f)
    for r in mine:
        e, f = _split(r)
        cands = idx.get(e)
        if not cands:
            return "row %r not in reference" % (r,)
        for i, g in enumerate(cands):
            if len(g) == len(f) and all(close(x, y, rtol) for x, y in zip(f, g)):
                cands.pop(i)
                break
        else:
            return "row %r: float fields differ from %r" % (r, cands[:2])
    return None
