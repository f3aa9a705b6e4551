"""TEST INFRASTRUCTURE ONLY -- runs the real reference build (oracle/_ref, see build_ref.py).

Imports the reference-generated CPython extension ``tpchref_sf<SF>_t<T>_compiled`` and calls
``<q>_compiled(db)`` exactly as the reference's dispatcher does (sdql_lib.py:410-424): ``db`` is a list
(one entry per query argument, in call order) of lists of numpy columns in schema order.

Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline / ``--impl reference`` legs may use
this module; the product path never does.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
MODS = os.path.join(HERE, "_ref", "mods")
SITE = os.path.join(HERE, "_ref", "site")

# argument order of each query = the benchmark() calls of test/test_all.py:1187-1208
QUERY_ARGS = {
    "q1": ["lineitem"], "q2": ["part", "supplier", "partsupp", "nation", "region"],
    "q3": ["lineitem", "customer", "orders"], "q4": ["orders", "lineitem"],
    "q5": ["lineitem", "customer", "orders", "region", "nation", "supplier"], "q6": ["lineitem"],
    "q7": ["supplier", "lineitem", "orders", "customer", "nation"],
    "q8": ["part", "supplier", "lineitem", "orders", "customer", "nation", "region"],
    "q9": ["lineitem", "orders", "nation", "supplier", "part", "partsupp"],
    "q10": ["customer", "orders", "lineitem", "nation"], "q11": ["partsupp", "supplier", "nation"],
    "q12": ["orders", "lineitem"], "q13": ["customer", "orders"], "q14": ["lineitem", "part"],
    "q15": ["lineitem", "supplier"], "q16": ["partsupp", "part", "supplier"], "q17": ["lineitem", "part"],
    "q18": ["lineitem", "customer", "orders"], "q19": ["lineitem", "part"],
    "q20": ["supplier", "nation", "partsupp", "part", "lineitem"],
    "q21": ["supplier", "lineitem", "orders", "nation"], "q22": ["customer", "orders"],
}


def available(name):
    import sysconfig
    return os.path.exists(os.path.join(MODS, name + "_compiled" + sysconfig.get_config_var("EXT_SUFFIX")))


def load(name):
    """name e.g. 'tpchref_sf1_t1' -> compiled module.  Needs oracle/_ref/{mods,site} (built by build_ref.py)."""
    for p in (MODS, SITE, os.path.join(SITE, "sdqlpy")):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module(name + "_compiled")


def normalise(res):
    """reference result -> float | int | sorted list of value tuples (record field order kept)."""
    if res is None or isinstance(res, (int, float)):
        return res
    d = res.to_dict().getContainer()  # fastd -> sr_dict{record: True} (fast_dict_generator.py:313-342)
    rows = []
    for rec in d.keys():
        rows.append(tuple(_py(v) for v in rec.getContainer().values()))
    return rows


def _py(v):
    if isinstance(v, str):
        return v.rstrip("\x00")
    return v


def run(mod, query, db):
    return normalise(getattr(mod, query + "_compiled")(db))


def check(mod, query, db, got, compare, tries=3):
    """compare ``got`` with the reference's result -> (difference or None, reference runs used).  The multi-threaded
    reference fills its ``dense`` bool sets (Q4 / Q21 / Q22: ``vector<bool>``, sdql_ir_cpp_generator_par.py:205, 738-741)
    from several threads without synchronisation, so neighbouring bits are occasionally lost and its own result changes
    from run to run; a mismatch is therefore re-checked against fresh reference runs."""
    d = None
    for k in range(tries):
        d = compare(got, run(mod, query, db))
        if d is None:
            return None, k + 1
    return d, tries
