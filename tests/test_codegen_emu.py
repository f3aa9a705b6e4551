"""generated query logic vs the REAL reference's outputs (tests/golden, produced by oracle/_ref).

The generated CUDA module is compiled as single-threaded host C++ (tests/emu, -DSDQLB200_EMU) so this runs in the
GPU-less dev container; it exercises the code generator, key packing, FD-minimised group keys, the C ABI and the
result boxing.  The same comparison runs against the real sm_100a build in test_gpu_parity.py."""
import os

import pytest

import emu
from compare import compare
from sdqlpy_b200 import build, runtime
from util import QUERIES, QUERY_SCRIPT, compact_db, golden, ref_db

import ref_runner as rr


@pytest.fixture(scope="module")
def emu_module(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu")
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py")
    cu = os.path.join(d, "q.cu")
    open(cu, "w").write(text)
    so = emu.build_emu(cu, os.path.join(d, "q_emu.so"))
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    yield runtime.CompiledModule(so)
    runtime.set_backend(old)
    runtime.STORE.clear()


@pytest.mark.parametrize("q", QUERIES)
def test_query_matches_reference_sf001(emu_module, q):
    gold = golden(0.01)
    got = emu_module.run(q, compact_db(0.01, rr.QUERY_ARGS[q]))
    assert compare(got, gold[q]) is None


@pytest.mark.parametrize("q", ["q1", "q3", "q6", "q10", "q13", "q16", "q22"])
def test_reference_layout_inputs(emu_module, q):
    """same queries fed with the reference's own input layout (int64 / float64 / <U n numpy arrays)."""
    gold = golden(0.01)
    got = emu_module.run(q, ref_db(0.01, rr.QUERY_ARGS[q]))
    assert compare(got, gold[q]) is None


def test_empty_relation(emu_module):
    import numpy as np
    from sdqlpy_b200.tpch.gen import SCHEMAS, Column
    cols = []
    for c, k in SCHEMAS["lineitem"]:
        if isinstance(k, tuple):
            cols.append(Column(c, "code", np.zeros(0, dtype=np.uint8), ["x"], k[1]))
        elif k == "float":
            cols.append(Column(c, "f64", np.zeros(0)))
        else:
            cols.append(Column(c, "i32", np.zeros(0, dtype=np.int32)))
    assert emu_module.run("q6", [cols]) == 0.0
    assert emu_module.run("q1", [cols]).size() == 0


FORCED = {"SDQLB200_COUNT_MIN_BYTES": "0", "SDQLB200_COUNT_MIN_RATIO": "0", "SDQLB200_BITS_MIN_BYTES": "0"}


@pytest.fixture(scope="module", params=["full_key_bits", "prefix_bits", "hash_tables"])
def forced_module(request, tmp_path_factory):
    """the same module with every optional table path switched on at tiny scale: cardinality passes in front of every
    selective build (right-sized tables), presence filters in front of every probed table (whole-key and first-part
    variants), and hash tables everywhere (no direct indexing).  The knobs are read once per loaded library."""
    env = dict(FORCED)
    if request.param == "prefix_bits":
        env["SDQLB200_BITS_PREFIX"] = "1"
    if request.param == "hash_tables":
        env["SDQLB200_FORCE_HASH"] = "1"
    old_env = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    d = tmp_path_factory.mktemp("emu_" + request.param)
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py")
    cu = os.path.join(d, "q.cu")
    open(cu, "w").write(text)
    so = emu.build_emu(cu, os.path.join(d, "q_emu_%s.so" % request.param))
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    yield runtime.CompiledModule(so)
    runtime.set_backend(old)
    runtime.STORE.clear()
    for k, v in old_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("q", QUERIES)
def test_forced_table_paths_match_reference(forced_module, q):
    gold = golden(0.01)
    got = forced_module.run(q, compact_db(0.01, rr.QUERY_ARGS[q]))
    assert compare(got, gold[q]) is None


@pytest.mark.parametrize("q", QUERIES)
def test_empty_relations_match_reference(emu_module, q):
    """every relation empty: the reference's outputs (tests/golden/tpch_empty.json: empty sets, 0.0, Q14's 0/0 = NaN,
    Q19's single all-zero record)"""
    from util import cut_db
    got = emu_module.run(q, cut_db(compact_db(0.01, rr.QUERY_ARGS[q]), rr.QUERY_ARGS[q], 0))
    assert compare(got, golden("empty")[q]) is None


@pytest.mark.parametrize("q", QUERIES)
def test_ragged_relations_match_reference(emu_module, q):
    """257 orders and their 1023 lineitems (no multiple of 4 / 32 / 128 rows), complete dimension tables"""
    import json
    from util import ROOT, cut_db
    rows = json.load(open(os.path.join(ROOT, "tests", "golden", "tpch_ragged.json")))["rows"]
    got = emu_module.run(q, cut_db(compact_db(0.01, rr.QUERY_ARGS[q]), rr.QUERY_ARGS[q], rows))
    assert compare(got, golden("ragged")[q]) is None


def test_code_generator_switches_keep_results(tmp_path):
    """the A/B partner builds of the defaults (64-bit row indices, register tier-0 accumulators, L2-prefetch pipeline,
    PROBE32 / RECONVERGE / TEXTSCAN / TEXTRESOLVE / TEXTALIGNED off) must reproduce the reference's outputs like the default build.  The code
    generator reads its switches at import, hence the subprocess."""
    import subprocess
    import sys
    script = r'''
import os, sys
sys.path[:0] = [%(root)r, %(root)r + "/tests", %(root)r + "/tests/emu", %(root)r + "/oracle"]
import emu
from compare import compare
from sdqlpy_b200 import build, runtime
from util import QUERY_SCRIPT, compact_db, golden
import ref_runner as rr
qs = ["q1", "q2", "q4", "q5", "q6", "q3", "q7", "q10", "q12", "q13", "q9", "q15", "q16", "q18", "q21", "q22"]
text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py", only=qs)
cu = os.path.join(%(tmp)r, os.environ["TAG"] + ".cu")
open(cu, "w").write(text)
runtime.set_backend(emu.EmuBackend())
mod = runtime.CompiledModule(emu.build_emu(cu, cu[:-3] + ".so"))
for q in qs:
    d = compare(mod.run(q, compact_db(0.01, rr.QUERY_ARGS[q])), golden(0.01)[q])
    assert d is None, (q, d)
print("ok")
''' % {"root": os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tmp": str(tmp_path)}
    for tag, env in (("idx64", {"SDQLB200_IDX32": "0", "SDQLB200_TEXTALIGNED": "0"}),
                     ("tier0reg", {"SDQLB200_TIER0_SMEM": "0", "SDQLB200_PIPELINE": "l2", "SDQLB200_TEXTRESOLVE": "0"}),
                     ("allreg", {"SDQLB200_PIPELINE": "reg"}),
                     ("plain", {"SDQLB200_PROBE32": "0", "SDQLB200_RECONVERGE": "0", "SDQLB200_TEXTSCAN": "0"})):
        e = dict(os.environ, TAG=tag, **env)
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=e)
        assert r.returncode == 0 and "ok" in r.stdout, (tag, r.stdout[-500:], r.stderr[-2000:])
