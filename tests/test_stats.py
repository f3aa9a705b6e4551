"""counting build (-DSDQLB200_STATS) and the bytes-moved figure, under the single-thread emulation (CPU)."""
import os

import pytest

import emu
from compare import compare
from sdqlpy_b200 import build, roofline, runtime
from util import QUERY_SCRIPT, compact_db, golden

import ref_runner as rr


@pytest.fixture(scope="module")
def stats_module(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu_stats")
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py", only=["q1", "q6", "q3", "q13", "q18"])
    cu = os.path.join(d, "q.cu")
    open(cu, "w").write(text)
    so = emu.build_emu(cu, os.path.join(d, "q_stats.so"), defines=("SDQLB200_STATS",))
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    yield runtime.CompiledModule(so)
    runtime.set_backend(old)
    runtime.STORE.clear()


def run(mod, q):
    db = compact_db(0.01, rr.QUERY_ARGS[q])
    mod.stats()  # reset
    got = mod.run(q, db)
    st, counting = mod.stats()
    assert counting
    assert compare(got, golden(0.01)[q]) is None  # counting does not change results
    man = mod.queries[q]
    nrows = {a: len(next(c for c in rel if c is not None).data) for a, rel in zip(man["args"], db)}
    return st, man, nrows


def test_scan_only_queries_have_no_random_accesses(stats_module):
    st, man, nrows = run(stats_module, "q6")
    assert st["finds"] == st["upserts"] == st["gathers"] == st["atomics"] == st["bit_tests"] == 0
    cols, strs = roofline.scan_bytes(man, nrows)
    assert (cols, strs) == (28 * nrows["li"], 0)
    bm = roofline.bytes_moved(man, nrows, st, 1)
    assert bm["total"] == cols + 8
    st, man, nrows = run(stats_module, "q1")
    assert roofline.scan_bytes(man, nrows) == (38 * nrows["li"], 0)


def test_join_query_counters(stats_module):
    st, man, nrows = run(stats_module, "q3")
    # every lineitem row that passes its date predicate looks its order up: probes happen, slots >= probes
    assert st["finds"] > 0 and st["find_slots"] >= st["finds"]
    assert st["upserts"] > 0 and st["upsert_slots"] >= st["upserts"]
    assert st["init_bytes"] > 0
    bm = roofline.bytes_moved(man, nrows, st, 10)
    assert bm["total"] > bm["scan_columns"] > 0 and bm["random_sectors"] % 32 == 0
    # a second call without a run in between reports nothing
    st2, _ = stats_module.stats()
    assert not any(st2.values())


def test_string_scan_bytes_are_counted(stats_module):
    st, man, nrows = run(stats_module, "q13")
    cols, strs = roofline.scan_bytes(man, nrows)
    assert strs == 79 * nrows["ord"]          # o_comment is searched for every order
    assert cols == 4 * nrows["ord"] + 4 * nrows["cu"]
    assert st["finds"] > 0                    # customers probe the per-customer order counts (the group-bys run in the
                                              # shared-memory tiers at this size: no global insert-or-find left to count)


def test_regular_build_reports_init_bytes_only():
    """the regular (timed) build exports sdqlb200_stats too, without device counters (checked on the emulation build)."""
    import tempfile
    d = tempfile.mkdtemp()
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py", only=["q3"])
    cu = os.path.join(d, "q.cu")
    open(cu, "w").write(text)
    so = emu.build_emu(cu, os.path.join(d, "q.so"))
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    try:
        mod = runtime.CompiledModule(so)
        mod.run("q3", compact_db(0.01, rr.QUERY_ARGS["q3"]))
        st, counting = mod.stats()
        assert not counting and st["finds"] == 0 and st["init_bytes"] > 0
    finally:
        runtime.set_backend(old)
        runtime.STORE.clear()
