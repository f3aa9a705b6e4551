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
