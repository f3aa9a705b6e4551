"""bench.py's result_check (size-independent properties of the bench query's result against numpy reductions over the
host columns) exercised on the emulation build: accepts the real result, rejects a corrupted one, never raises."""
import os
import sys

import pytest

import emu
from sdqlpy_b200 import build, runtime
from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH
from util import QUERY_SCRIPT, ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu_bench")
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py", only=["q1", "q6"])
    cu = os.path.join(d, "q.cu")
    open(cu, "w").write(text)
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    mod = runtime.CompiledModule(emu.build_emu(cu, os.path.join(d, "q.so")))
    cols = TPCH(0.01).columns("lineitem")
    yield mod, cols, [[cols.get(c) for c, _ in SCHEMAS["lineitem"]]]
    runtime.set_backend(old)
    runtime.STORE.clear()


def test_accepts_the_real_results(setup):
    mod, cols, db = setup
    assert bench.result_check("q1", cols, mod.run("q1", db)) == "ok"
    assert bench.result_check("q6", cols, mod.run("q6", db)) == "ok"


def test_rejects_corrupted_results_without_raising(setup):
    mod, cols, db = setup
    r = mod.run("q1", db)
    rows = r.tuples()
    bad = runtime.ResultSet(r.names, [rows[0][:-1] + (rows[0][-1] + 1,)] + rows[1:])
    assert "count_order" in bench.result_check("q1", cols, bad)
    bad = runtime.ResultSet(r.names, rows[1:])
    assert bench.result_check("q1", cols, bad) != "ok"
    assert "revenue" in bench.result_check("q6", cols, mod.run("q6", db) * (1 + 1e-6))
    assert bench.result_check("q6", {}, 1.0).startswith("checker failed")
    assert bench.result_check("q3", cols, None).startswith("no check")
