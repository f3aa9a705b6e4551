"""the reference-facing front door: sdqlpy_init(mode, N) / @sdql_compile dispatch / benchmark(), the importable
<script>_compiled module the UNMODIFIED reference wrapper imports (sdqlpy/sdql_lib.py:401-424), the fastd surface
(fastd.py:31-51, fast_dict_generator.py:241-342) and the N-GPU engine's host logic -- on the CPU through the emulation
build (tests/emu); the same entry points run on the B200 in test_gpu_front_door.py."""
import importlib
import os
import sys

import numpy as np
import pytest

import emu
import ref_runner as rr
from compare import compare
from sdqlpy_b200 import build, runtime, sdql_lib, tbl
from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH
from util import QUERY_SCRIPT, ROOT, compact_db, golden, ref_db

QS = ["q1", "q6", "q3", "q13"]


def script_text(lib_import):
    """a user script in the reference's style: the schemas and four of the 22 queries, cut out of the workload file"""
    src = open(QUERY_SCRIPT).read()
    head = src[src.index("lineitem_type = "):src.index("@sdql_compile")]
    out = [lib_import, head]
    for q in QS:
        i = src.index("@sdql_compile", src.index("def %s(" % q) - 120)
        j = src.find("\n@sdql_compile", i + 10)
        out.append(src[i:j if j > 0 else len(src)].rstrip() + "\n\n")
    return "\n".join(out)


@pytest.fixture(scope="module")
def user_script(tmp_path_factory):
    """<tmp>/myq.py + its emulation-built module at the place build.out_paths names + the <script>_compiled stub"""
    d = str(tmp_path_factory.mktemp("front"))
    path = os.path.join(d, "myq.py")
    open(path, "w").write(script_text("from sdqlpy.sdql_lib import *"))
    text, _ = build.compile_source(script_text("from sdqlpy_b200.sdql_lib import *"), "myq.py")
    cu, so = build.out_paths(path)
    os.makedirs(os.path.dirname(cu), exist_ok=True)
    open(cu, "w").write(text)
    emu.build_emu(cu, so)
    build.write_stub(path)
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    yield path
    runtime.set_backend(old)
    runtime.STORE.clear()
    sys.modules.pop("myq_compiled", None)


def write_tbl(d, table, sf=0.01):
    g = TPCH(sf)
    p = os.path.join(d, table + ".tbl")
    open(p, "wb").write(tbl.format_tbl(SCHEMAS[table], g.ref_table(table, [c for c, _ in SCHEMAS[table]])))
    return p


def test_unmodified_reference_wrapper_dispatches_to_the_stub(user_script, monkeypatch):
    """the reference's own sdql_lib (scratch copy under oracle/_ref/site, sdql_lib.py untouched) in mode 2: its decorator
    imports myq_compiled, finds q<N>_compiled and calls it with db = [arg["data"], ..]; the result is a ``fastd``"""
    site = os.path.join(ROOT, "oracle", "_ref", "site")
    if not os.path.isdir(os.path.join(site, "sdqlpy")):
        pytest.skip("oracle/_ref not built")
    d = os.path.dirname(user_script)
    monkeypatch.syspath_prepend(site)
    monkeypatch.syspath_prepend(os.path.join(site, "sdqlpy"))
    monkeypatch.syspath_prepend(d)
    ref = importlib.import_module("sdqlpy.sdql_lib")
    ref.sdqlpy_init(2, 1)
    myq = importlib.import_module("myq")
    li = ref.read_csv(write_tbl(d, "lineitem"), myq.lineitem_type, "li")
    cu = ref.read_csv(write_tbl(d, "customer"), myq.customer_type, "cu")
    od = ref.read_csv(write_tbl(d, "orders"), myq.order_type, "ord")
    gold = golden(0.01)
    assert compare(myq.q6(li), gold["q6"]) is None
    r1 = myq.q1(li)
    assert r1.__class__.__name__ == "fastd" and r1.size() == len(gold["q1"])
    assert compare(r1, gold["q1"]) is None
    dd = r1.to_dict()
    assert isinstance(dd, ref.sr_dict) and all(isinstance(k, ref.record) and v is True for k, v in dd.getContainer().items())
    assert list(next(iter(dd.getContainer())).getContainer().keys()) == [
        "l_returnflag", "l_linestatus", "sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "count_order"]
    assert compare(myq.q3(li, cu, od), gold["q3"]) is None
    assert compare(myq.q13(cu, od), gold["q13"]) is None
    sys.modules.pop("myq", None)


def test_own_front_door_modes_and_benchmark(user_script, capsys, monkeypatch):
    """sdqlpy_b200.sdql_lib: mode 2 dispatch through @sdql_compile, benchmark() protocol and printing (lib:437-475)"""
    d = os.path.dirname(user_script)
    path = os.path.join(d, "ownq.py")
    open(path, "w").write(script_text("from sdqlpy_b200.sdql_lib import *"))
    cu, so = build.out_paths(path)
    os.makedirs(os.path.dirname(cu), exist_ok=True)
    import shutil
    shutil.copy(build.out_paths(user_script)[1], so)
    monkeypatch.syspath_prepend(d)
    ownq = importlib.import_module("ownq")
    li = sdql_lib.table_from_columns([c for c, _ in SCHEMAS["lineitem"]], ref_db(0.01, ["lineitem"])[0])
    sdql_lib.sdqlpy_init(2, 1)
    try:
        assert compare(ownq.q6(li), golden(0.01)["q6"]) is None
        times = sdql_lib.benchmark("Q1", 2, ownq.q1, [li])
        out = capsys.readouterr().out
        assert len(times) == 2 and "Q1: Mean:" in out and "Result Size: %d" % len(golden(0.01)["q1"]) in out
        assert "{ <" in out and "> -> true" in out   # the reference's print format (phmap.h:78-93)
        sdql_lib.sdqlpy_init(7, 1)
        assert "not supported" in capsys.readouterr().out
    finally:
        sdql_lib.sdqlpy_init(0, 1)
        sys.modules.pop("ownq", None)


def test_fastd_surface(user_script):
    """size / print / to_dict / get / set / from_dict as the reference's FastDict type behaves (fdg:241-342)"""
    mod = runtime.load_compiled(user_script)
    res = mod.run("q1", compact_db(0.01, ["lineitem"]))
    gold = golden(0.01)["q1"]
    n = res.size()
    assert n == len(gold) == len(res)
    text = str(res)
    assert text.startswith("{ <") and text.endswith("> -> true }") and text.count(" -> true") == n
    row = res.tuples()[0]
    assert ("%.2f" % row[2]) in text  # fixed 2-decimal floats
    d = res.to_dict()
    assert isinstance(d, sdql_lib.sr_dict) and len(d) == n
    key = next(iter(d.getContainer()))
    assert res.get(key) is True and res.size() == n
    missing = sdql_lib.record(dict(zip(res.names, ("X", "Y", 1.0, 2.0, 3.0, 4.0, 5))))
    assert res.get(missing) is False and res.size() == n + 1       # operator[] inserts the key with false (fdg:289)
    assert "<X,Y,1.00,2.00,3.00,4.00,5> -> false" in str(res)
    assert res.set(missing, True) is True and res.get(missing) is True and res.size() == n + 1
    other = sdql_lib.record(dict(zip(res.names, ("Z", "W", 0.5, 0.5, 0.5, 0.5, 1))))
    assert res.from_dict({other: True}) is True and res.size() == n + 2
    assert len(res.to_dict()) == n + 2
    empty = mod.run("q1", [[None if c is None else type(c)(c.name, c.kind, c.data[:0], c.dictionary, c.width) for c in compact_db(0.01, ["lineitem"])[0]]])
    assert empty.size() == 0 and str(empty) == "{  }"


def test_engine_partitions_on_key_boundaries():
    g = TPCH(0.01)
    eng = runtime.Engine(3, _test_backend=emu.EmuBackend)
    try:
        q = {"args": ["li", "cu", "ord"], "schemas": {"li": SCHEMAS["lineitem"], "cu": SCHEMAS["customer"], "ord": SCHEMAS["orders"]}}
        db = compact_db(0.01, ["lineitem", "customer", "orders"])
        dbs, parts = eng.partition(q, db)
        assert parts == {"li", "ord"}
        nli = sum(len(d[0][0].data) for d in dbs)
        assert nli == len(db[0][0].data) and sum(len(d[2][0].data) for d in dbs) == len(db[2][0].data)
        for d in dbs:   # co-partitioned: a rank's lineitems belong to its orders
            assert set(np.unique(d[0][0].data)) <= set(d[2][0].data.tolist())
            assert d[1] is db[1]  # customer replicated
        assert eng.partition(q, db)[0][1][0][0] is dbs[1][0][0]  # slices are cached: stable identity for the column store
        eng2 = runtime.Engine(2, partitioned=None, _test_backend=emu.EmuBackend)
        dbs2, parts2 = eng2.partition(q, db)
        assert parts2 == {"li"} and [len(d[0][0].data) for d in dbs2] == [len(db[0][0].data) // 2, len(db[0][0].data) - len(db[0][0].data) // 2]
        eng2.close()
    finally:
        eng.close()


@pytest.mark.parametrize("world", [2, 3])
def test_sdqlpy_init_n_runs_on_n_ranks(user_script, world, monkeypatch):
    """sdqlpy_init(2, N): N ranks in this process (threads; here each with the emulation back end), lineitem / orders
    range partitioned, partial results merged and concatenated -- same answers as on one rank"""
    d = os.path.dirname(user_script)
    path = os.path.join(d, "engq%d.py" % world)
    open(path, "w").write(script_text("from sdqlpy_b200.sdql_lib import *"))
    cu, so = build.out_paths(path)
    os.makedirs(os.path.dirname(cu), exist_ok=True)
    import shutil
    shutil.copy(build.out_paths(user_script)[1], so)
    monkeypatch.syspath_prepend(d)
    engq = importlib.import_module("engq%d" % world)
    tabs = {t: sdql_lib.table_from_columns([c for c, _ in SCHEMAS[t]], compact_db(0.01, [t])[0]) for t in ("lineitem", "customer", "orders")}
    sdql_lib.sdqlpy_init(2, world)
    sdql_lib._state["engine"] = runtime.Engine(world, _test_backend=emu.EmuBackend)
    try:
        gold = golden(0.01)
        assert compare(engq.q6(tabs["lineitem"]), gold["q6"]) is None
        assert compare(engq.q1(tabs["lineitem"]), gold["q1"]) is None
        assert compare(engq.q3(tabs["lineitem"], tabs["customer"], tabs["orders"]), gold["q3"]) is None
        assert compare(engq.q13(tabs["customer"], tabs["orders"]), gold["q13"]) is None
        assert compare(engq.q1(tabs["lineitem"]), gold["q1"]) is None  # second call: cached slices / statistics
    finally:
        sdql_lib.sdqlpy_init(0, 1)
        sys.modules.pop("engq%d" % world, None)


def test_partitioned_string_dictionaries_agree(user_script):
    """reference-layout (<U n) string columns of a partitioned relation: every rank dictionary-encodes its own rows, the
    ranks then agree on the union (codes key the merged tables).  23 lineitems over 3 ranks: no rank sees every flag."""
    if not rr.available("tpchref_sf1_t1"):
        pytest.skip("oracle/_ref not built")
    from util import cut_db
    mod = runtime.load_compiled(user_script)
    db = cut_db(ref_db(0.01, ["lineitem"]), ["lineitem"], 23)
    flags = [set(db[0][8][lo:hi].tolist()) for lo, hi in ((0, 7), (7, 15), (15, 23))]
    assert len(set(map(frozenset, flags))) > 1  # the ranks' local dictionaries differ
    want = rr.run(rr.load("tpchref_sf1_t1"), "q1", db)
    eng = runtime.Engine(3, partitioned=("li",), partkeys=(), _test_backend=emu.EmuBackend)
    try:
        assert compare(eng.run(mod, "q1", db), want) is None
    finally:
        eng.close()


def test_stride_statistic_is_verified_on_every_value():
    """sdqlb200_col.stride: dbgen order keys use 8 of every 32 values -> (log2 32 << 16) | 8; one key that breaks the rule
    anywhere in the column (missed by the strided sample) must cancel it, or the packed tables would drop that key"""
    g = TPCH(0.05)
    ok = g.columns("orders", ["o_orderkey"])["o_orderkey"].data
    assert runtime.stride_stat(ok, int(ok.min())) == (5 << 16) | 8
    bad = ok.copy()
    bad[len(bad) // 2 + 1] += 9            # residue 9..16: outside the first 8 of its block
    assert runtime.stride_stat(bad, int(bad.min())) in (0, (5 << 16) | 16, (5 << 16) | 15, (5 << 16) | 14, (5 << 16) | 13, (5 << 16) | 12,
                                                         (5 << 16) | 11, (5 << 16) | 10)
    assert runtime.stride_stat(bad, int(bad.min())) != (5 << 16) | 8
    dense = g.columns("customer", ["c_custkey"])["c_custkey"].data
    assert runtime.stride_stat(dense, int(dense.min())) == 0
    assert runtime.stride_stat(ok[:100], int(ok.min())) == 0   # too few rows to bother


def test_column_store_invalidate_forgets_one_host_column():
    """ColumnStore contract: host arrays are immutable while cached; after an in-place update STORE.invalidate(array) drops
    the device copies of exactly that column (every representation)"""
    import numpy as np
    st = runtime.ColumnStore()
    a, b = np.arange(8, dtype=np.int64), np.arange(8, dtype=np.float64)
    st.cache[(st.key(a), "i32", 0)] = ("dev_a", a)
    st.cache[(st.key(a), "f64", 0)] = ("dev_a64", a)
    st.cache[(st.key(b), "f64", 0)] = ("dev_b", b)
    st.invalidate(a)
    assert list(st.cache) == [(st.key(b), "f64", 0)]
