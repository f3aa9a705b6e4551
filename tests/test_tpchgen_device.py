"""device-side TPC-H generator (csrc/sdqlb200_tpchgen.cu) vs the host generator gen.TPCH: bit-identical columns for
any order range, and queries fed with device-generated fact tables match the reference's golden outputs."""
import ctypes
import os
import re

import numpy as np
import pytest

from sdqlpy_b200 import build
from sdqlpy_b200.tpch import gen
from util import ROOT


def test_library_exports_declared_symbols():
    lib = ctypes.CDLL(build.compile_tpchgen())
    hdr = open(os.path.join(ROOT, "include", "sdqlb200_tpchgen.h")).read()
    names = set(re.findall(r"\b(sdqlb200_tpchgen_[a-z_]+)\s*\(", hdr))
    assert names >= {"sdqlb200_tpchgen_order_lines", "sdqlb200_tpchgen_lineitem", "sdqlb200_tpchgen_orders",
                     "sdqlb200_tpchgen_orders_text", "sdqlb200_tpchgen_last_error"}
    for n in names:
        assert hasattr(lib, n), n


def _host(dc):
    import torch
    t = dc.holder
    a = t.cpu().numpy()
    return a[:dc.rows]


@pytest.mark.gpu
@pytest.mark.parametrize("rng", [None, (100, 1777), (0, 1), (74999, 75000)])
def test_columns_bit_identical_to_host_generator(rng):
    from sdqlpy_b200 import runtime
    from sdqlpy_b200.tpch.gen_device import DeviceTPCH
    runtime.set_backend(None)
    sf = 0.05
    h, d = gen.TPCH(sf), DeviceTPCH(sf)
    assert d.O == h.O == 75000
    for table in ("lineitem", "orders"):
        hc = h.columns(table, None, rng)
        dc = d.columns(table, None, rng)
        assert set(dc) == set(hc), (sorted(dc), sorted(hc))
        for name, col in hc.items():
            got = _host(dc[name])
            want = col.data
            assert got.shape == want.shape, (name, got.shape, want.shape)
            if want.dtype == np.float64:
                assert (got.view(np.int64) == want.view(np.int64)).all(), name
            else:
                assert (got == want).all(), name
            if col.kind == "i32" and len(want):
                assert (dc[name].min, dc[name].max) == (int(want.min()), int(want.max()))
            if col.kind == "code":
                assert dc[name].dictionary == list(col.dictionary)
    assert d.rows("lineitem", rng) == len(next(iter(h.columns("lineitem", ["l_orderkey"], rng).values())).data)


@pytest.mark.gpu
@pytest.mark.parametrize("q", ["q1", "q3", "q4", "q6", "q10", "q12", "q13", "q18", "q21"])
def test_queries_on_device_generated_fact_tables(q):
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import runtime
    from sdqlpy_b200.tpch.gen_device import DeviceTPCH
    from util import QUERY_SCRIPT, compact_db, golden
    runtime.set_backend(None)
    mod = runtime.load_compiled(QUERY_SCRIPT)
    d = DeviceTPCH(0.05)
    db = []
    for t, host_rel in zip(rr.QUERY_ARGS[q], compact_db(0.05, rr.QUERY_ARGS[q])):
        if t in ("lineitem", "orders"):
            cols = d.columns(t)
            db.append([cols.get(c) for c, _ in gen.SCHEMAS[t]])
        else:
            db.append(host_rel)
    assert compare(mod.run(q, db), golden(0.05)[q]) is None


@pytest.mark.gpu
def test_bad_order_range_is_an_error():
    from sdqlpy_b200 import runtime
    from sdqlpy_b200.tpch.gen_device import DeviceTPCH
    runtime.set_backend(None)
    d = DeviceTPCH(0.01)
    with pytest.raises(RuntimeError, match="order range"):
        d._ck(__import__("sdqlpy_b200.tpch.gen_device", fromlist=["lib"]).lib().sdqlb200_tpchgen_order_lines(
            ctypes.byref(d.params), 0, d.O + 5, None, None), "tpchgen_order_lines")
