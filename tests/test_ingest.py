"""device-side ingest of reference-layout columns (csrc/sdqlb200_ingest.cu) against the numpy restatement the CPU path of
the tests uses (runtime._encode): int64 -> int32 + statistics, <U n -> bytes, <U n -> dictionary codes; and the exchange /
ingest libraries export what their headers declare."""
import ctypes
import os
import re

import numpy as np
import pytest

from sdqlpy_b200 import build, runtime
from util import ROOT, ref_db


@pytest.mark.parametrize("name,fn", [("sdqlb200_ingest", "compile_ingest"), ("sdqlb200_comm", "compile_comm")])
def test_library_exports_declared_symbols(name, fn):
    so = getattr(build, fn)()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", name + ".h")).read()
    names = set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % name, hdr))
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), n


@pytest.fixture()
def cuda_be():
    runtime.set_backend(None)
    return runtime.backend()


def _host(holder, n, dtype):
    return holder.reshape(-1)[:n].cpu().numpy().astype(dtype) if holder.dim() == 1 else holder[:n].cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [1 << 29, 4096])
def test_ingest_matches_host_encoding(cuda_be, chunk, monkeypatch):
    from sdqlpy_b200 import ingest
    monkeypatch.setattr(ingest, "CHUNK_BYTES", chunk)
    li = ref_db(0.01, ["lineitem"])[0]
    cu = ref_db(0.01, ["customer"])[0]
    for a in (li[0], li[10], cu[3]):  # l_orderkey, l_shipdate, c_nationkey (int64)
        img, mn, mx, w, _ = runtime._encode(a, "i32", 0)
        ptr, holder, gmn, gmx, gw, _, h2d = ingest.upload(a, "i32", 0, cuda_be)
        assert (gmn, gmx, gw, h2d) == (mn, mx, 4, a.nbytes)
        assert np.array_equal(_host(holder, len(a), np.int32), img)
    for a, width in ((li[13], 25), (cu[4], 15), (cu[7], 117)):  # l_shipinstruct, c_phone, c_comment -> bytes
        img, _, _, w, _ = runtime._encode(a, "bytes", width)
        ptr, holder, _, _, gw, _, _ = ingest.upload(a, "bytes", width, cuda_be)
        assert gw == width and np.array_equal(holder[:len(a)].cpu().numpy(), img)
    for a in (li[8], li[13], li[14], cu[6], cu[1]):  # flags, instructions, modes, segments, c_name (1500 distinct -> int32 codes)
        img, _, mx, w, d = runtime._encode(a, "code", 0)
        ptr, holder, _, gmx, gw, gd, _ = ingest.upload(a, "code", 0, cuda_be)
        assert gd == d and gmx == mx and gw == w
        assert np.array_equal(_host(holder, len(a), img.dtype), img)
    empty = ingest.upload(np.zeros(0, dtype=np.int64), "i32", 0, cuda_be)
    assert empty[2:5] == (0, 0, 4)


@pytest.mark.gpu
def test_ingest_errors(cuda_be, monkeypatch):
    from sdqlpy_b200 import ingest
    with pytest.raises(ValueError, match="int32 range"):
        ingest.upload(np.array([1, 2**40, 3], dtype=np.int64), "i32", 0, cuda_be)
    with pytest.raises(ValueError, match="non-latin1"):
        ingest.upload(np.array(["ab", "cЖ"], dtype="<U4"), "bytes", 4, cuda_be)
    monkeypatch.setattr(ingest, "DICT_SLOTS", 64)
    many = np.array(["v%d" % i for i in range(100)], dtype="<U6")
    with pytest.raises(ingest.TooManyValues):
        ingest.upload(many, "code", 0, cuda_be)
    # the column store falls back to the host dictionary for such a column
    runtime.STORE.clear()
    col = runtime.STORE.get(many, "code", 0)
    assert col.dictionary == sorted(many.tolist()) and col.width == 1
    runtime.STORE.clear()


def test_host_narrowing_passes():
    """the host side of the upload (sdqlb200_ingest_host_*: plain C++ threads in the ingest library, no device involved):
    int64 -> int32 with min / max, <U1 -> bytes with the set of bytes present and the first row that does not fit a byte"""
    lib = ctypes.CDLL(build.compile_ingest())
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    lib.sdqlb200_ingest_host_i64.argtypes = [vp, vp, i64, i32, vp]
    lib.sdqlb200_ingest_host_ucs4_1.argtypes = [vp, vp, i64, i32, vp, vp]
    rng = np.random.default_rng(7)
    for n in (0, 1, 63, 64, 65, 1000, 100003):
        a = rng.integers(-2**31, 2**31, n, dtype=np.int64)
        out, mm = np.full(max(n, 1), 123, dtype=np.int32), np.zeros(2, dtype=np.int64)
        for th in (1, 3, 8):
            assert lib.sdqlb200_ingest_host_i64(a.ctypes.data, out.ctypes.data, n, th, mm.ctypes.data) == 0
            assert np.array_equal(out[:n], a.astype(np.int32))
            if n:
                assert (int(mm[0]), int(mm[1])) == (int(a.min()), int(a.max()))
        u = rng.choice(np.array(["A", "N", "R", ""], dtype="<U1"), n)
        o8, pr, bad = np.full(max(n, 1), 9, dtype=np.uint8), np.zeros(4, dtype=np.uint64), np.zeros(1, dtype=np.int64)
        for th in (1, 5):
            assert lib.sdqlb200_ingest_host_ucs4_1(u.ctypes.data, o8.ctypes.data, n, th, pr.ctypes.data, bad.ctypes.data) == 0
            assert np.array_equal(o8[:n], u.view(np.uint32).astype(np.uint8)) and int(bad[0]) == -1
            present = {b for b in range(256) if (int(pr[b >> 6]) >> (b & 63)) & 1}
            assert present == {int(x) for x in np.unique(u.view(np.uint32))}
        if n > 10:
            u[n // 2], u[n - 1] = "Ж", "€"
            lib.sdqlb200_ingest_host_ucs4_1(u.ctypes.data, o8.ctypes.data, n, 4, pr.ctypes.data, bad.ctypes.data)
            assert int(bad[0]) == n // 2


@pytest.mark.gpu
def test_host_narrowed_upload_equals_device_ingest(cuda_be, monkeypatch):
    """SDQLB200_HOST_NARROW path of ColumnStore.get_many: the same resident columns, statistics and dictionaries as the
    device-side ingest, with fewer bytes over the link; a query through it matches the golden"""
    from compare import compare
    from util import QUERY_SCRIPT, golden
    import ref_runner as rr
    monkeypatch.setattr(runtime, "HOST_NARROW", True)
    monkeypatch.setattr(runtime, "HOST_NARROW_MIN_ROWS", 1)
    li = ref_db(0.05, ["lineitem"])[0]
    runtime.STORE.clear()
    items = [(li[10], "i32", 0, None), (li[4], "f64", 0, None), (li[8], "code", 1, None), (li[9], "code", 1, None), (li[0], "i32", 0, None)]
    h0 = runtime.STORE.h2d_bytes
    got = runtime.STORE.get_many(items)
    narrowed = runtime.STORE.h2d_bytes - h0
    runtime.STORE.clear()
    monkeypatch.setattr(runtime, "HOST_NARROW", False)
    h0 = runtime.STORE.h2d_bytes
    want = runtime.STORE.get_many(items)
    raw = runtime.STORE.h2d_bytes - h0
    n = len(li[0])
    assert narrowed == n * (4 + 8 + 1 + 1 + 4) and raw == n * (8 + 8 + 4 + 4 + 8)
    for g, w in zip(got, want):
        assert (g.kind, g.rows, g.min, g.max, g.width, g.dictionary, g.stride) == (w.kind, w.rows, w.min, w.max, w.width, w.dictionary, w.stride)
        assert np.array_equal(g.holder[:n].cpu().numpy(), w.holder[:n].cpu().numpy())
    runtime.STORE.clear()
    monkeypatch.setattr(runtime, "HOST_NARROW", True)
    mod = runtime.load_compiled(QUERY_SCRIPT)
    for q in ("q1", "q12"):
        assert compare(mod.run(q, ref_db(0.05, rr.QUERY_ARGS[q])), golden(0.05)[q]) is None
    # a <U1 column with a code point that does not fit a byte takes the device-side encoder
    odd = np.array(["A", "Ж", "A", "B"] * 10, dtype="<U1")
    col = runtime.STORE.get_many([(odd, "code", 1, None)])[0]
    assert col.dictionary == ["A", "B", "Ж"]
    runtime.STORE.clear()
