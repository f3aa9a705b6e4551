"""column wire formats: decode(encode(x)) == x bit for bit.
CPU: the numpy encoder / restated decoder on generator columns and adversarial columns, the library's exported
symbols, and query parity when every column crosses the link packed (emulation build).
GPU: the CUDA decode kernels (through the C ABI of include/sdqlb200_wire.h) against the original columns."""
import ctypes
import os
import re

import numpy as np
import pytest

from sdqlpy_b200 import build, wire
from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH
from util import ROOT


def bits(a):
    return a.view(np.int64) if a.dtype == np.float64 else a


def test_generator_columns_round_trip():
    g = TPCH(0.05)
    kinds = {}
    for t in ("lineitem", "orders", "customer", "partsupp", "part", "supplier"):
        for name, col in g.columns(t).items():
            if col.kind not in ("i32", "f64"):
                continue
            p = wire.pack(col.data, col.kind)
            if p is None:
                continue
            kinds[name] = wire.KIND_NAMES[p.kind]
            out = p.decode_host()
            assert out.dtype == col.data.dtype and (bits(out) == bits(col.data)).all(), name
            assert p.codes.nbytes < col.data.nbytes
            if col.kind == "i32":
                assert (p.min, p.max) == (int(col.data.min()), int(col.data.max()))
    # the Q1 / Q6 columns travel narrow: 53 bits instead of 38 bytes per lineitem row
    assert kinds["l_quantity"] == "bits_dict_f64" and kinds["l_discount"] == "bits_dict_f64" and kinds["l_tax"] == "bits_dict_f64"
    assert kinds["l_shipdate"] == "bits_dict_i32" and kinds["l_extendedprice"] == "bits_fixed_f64"
    assert kinds["o_totalprice"] == "bits_fixed_f64" and kinds["c_acctbal"] in ("bits_dict_f64", "bits_fixed_f64", "dict16_f64")
    li = g.columns("lineitem")
    q1 = {}
    for c in ("l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate", "l_returnflag", "l_linestatus"):
        wire.pack_column(li[c])
        q1[c] = li[c].wire.nbits
    assert q1 == {"l_quantity": 6, "l_extendedprice": 24, "l_discount": 4, "l_tax": 4, "l_shipdate": 12,
                  "l_returnflag": 2, "l_linestatus": 1}, q1
    for c in ("l_returnflag", "l_linestatus"):
        assert (li[c].wire.decode_host() == li[c].data).all()


def test_byte_aligned_kinds_without_bit_packing(monkeypatch):
    monkeypatch.setattr(wire, "BITPACK", False)
    g = TPCH(0.05)
    li = g.columns("lineitem", ["l_quantity", "l_shipdate", "l_extendedprice"])
    kinds = {n: wire.KIND_NAMES[wire.pack(c.data, c.kind).kind] for n, c in li.items()}
    assert kinds == {"l_quantity": "dict8_f64", "l_shipdate": "dict16_i32", "l_extendedprice": "fixed32_f64"}


@pytest.mark.parametrize("nbits", [1, 2, 3, 7, 8, 12, 13, 24, 31, 32])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1025, (1 << 20) + 3])
def test_bit_stream_round_trip(nbits, n):
    rng = np.random.default_rng(nbits * 1000 + n)
    c = rng.integers(0, 2**nbits, n, dtype=np.uint64)
    s = wire.bitpack(c, nbits)
    assert s.dtype == np.uint8 and len(s) % 16 == 0 and len(s) >= (n * nbits + 7) // 8 + 8
    assert (wire.bitunpack(s, nbits, n) == c).all()
    if n:  # field i sits at bit i * nbits, little endian
        i = n - 1
        word = int.from_bytes(bytes(s[(i * nbits) // 8:(i * nbits) // 8 + 8]), "little")
        assert (word >> ((i * nbits) % 8)) & (2**nbits - 1) == int(c[i])


@pytest.mark.parametrize("case", ["neg_zero", "nan", "wide", "not_cents", "many_ints", "late_values"])
def test_adversarial_columns_never_lose_bits(case):
    rng = np.random.default_rng(7)
    n = 5000
    if case == "neg_zero":
        a = np.where(rng.integers(0, 2, n) == 0, 0.0, -0.0)
    elif case == "nan":
        a = rng.integers(0, 5, n).astype(np.float64)
        a[17] = np.nan
    elif case == "wide":  # cents that do not fit int32 and > 65536 distinct values
        a = rng.integers(0, 2**40, 3 * 2**20).astype(np.float64) / 100.0
    elif case == "not_cents":
        a = rng.random(3 * 2**20)
    elif case == "many_ints":
        a = rng.integers(-2**31, 2**31 - 1, 3 * 2**20).astype(np.int32)
    else:  # values that only appear after the sampled prefix
        a = np.zeros(wire.SAMPLE + 4096, dtype=np.float64)
        a[wire.SAMPLE + 5:] = 123.25
    rep = "i32" if a.dtype == np.int32 else "f64"
    p = wire.pack(a, rep)
    if p is not None:
        out = p.decode_host()
        assert (bits(out) == bits(a)).all()
    if case in ("wide", "not_cents", "many_ints", "neg_zero"):  # (a single NaN payload packs exactly)
        assert p is None
    if case == "late_values":
        assert p is not None and p.kind == wire.BITS_DICT_F64 and p.nbits == 1


def test_negative_money_is_exact():
    cents = np.arange(-99999, 999999, 7, dtype=np.int64)
    a = cents / 100.0
    p = wire.pack(np.tile(a, 1), "f64")
    assert p is not None and p.kind == wire.BITS_FIXED_F64 and p.base == -99999
    assert (bits(p.decode_host()) == bits(a)).all()


def test_library_exports_declared_symbols():
    so = build.compile_wire()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "sdqlb200_wire.h")).read()
    names = set(re.findall(r"\b(sdqlb200_wire_[a-z_0-9]+)\s*\(", hdr))
    assert names >= {"sdqlb200_wire_decode", "sdqlb200_wire_decode_bits", "sdqlb200_wire_src_width",
                     "sdqlb200_wire_dst_width", "sdqlb200_wire_last_error"}
    for n in names:
        assert hasattr(lib, n), n
    for k in range(10):
        assert lib.sdqlb200_wire_src_width(k) == wire.SRC_WIDTH[k]
        assert lib.sdqlb200_wire_dst_width(k) == np.dtype(wire.DST_DTYPE[k]).itemsize
    assert lib.sdqlb200_wire_src_width(99) == 0


def packed_db(sf, tables):
    g = TPCH(sf)
    db = []
    for t in tables:
        cols = g.columns(t)
        db.append([cols.get(c) for c, _ in SCHEMAS[t]])
    wire.pack_db(db)
    return db


@pytest.mark.parametrize("q", ["q1", "q6", "q3", "q14", "q18"])
def test_queries_on_packed_columns_emu(q, tmp_path_factory):
    """host logic of the packed upload path (ColumnStore -> backend.upload_packed) under the emulation build."""
    import emu
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import runtime
    from util import QUERY_SCRIPT, golden
    d = tmp_path_factory.mktemp("emu_wire")
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py", only=[q])
    cu = os.path.join(d, "q.cu")
    open(cu, "w").write(text)
    so = emu.build_emu(cu, os.path.join(d, "q_emu.so"))
    old = runtime._backend
    runtime.set_backend(emu.EmuBackend())
    runtime.STORE.clear()
    try:
        mod = runtime.CompiledModule(so)
        db = packed_db(0.05, rr.QUERY_ARGS[q])
        assert any(c is not None and c.wire is not None for rel in db for c in rel)
        assert compare(mod.run(q, db), golden(0.05)[q]) is None
        assert 0 < mod.last.h2d_bytes
    finally:
        runtime.set_backend(old)
        runtime.STORE.clear()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("rows", [0, 1, 3, 1024, 1027, 1 << 20, 3_000_001])
def test_device_decode_bit_exact(rows, monkeypatch):
    import torch
    monkeypatch.setattr(wire, "BITPACK", False)  # the byte-aligned kinds (the bit-packed ones have their own test)
    from sdqlpy_b200 import runtime
    runtime.set_backend(None)
    be = runtime.backend()
    rng = np.random.default_rng(rows)
    cols = {
        wire.DICT8_F64: rng.integers(0, 11, rows) / 100.0,
        wire.DICT16_F64: rng.integers(0, 9999, rows).astype(np.float64),
        wire.DICT8_I32: rng.integers(1, 51, rows).astype(np.int32),
        wire.DICT16_I32: (19920101 + rng.integers(0, 2500, rows)).astype(np.int32),
        wire.FIXED32_F64: rng.integers(-99999, 10_500_000, rows) / 100.0,
    }
    for kind, a in cols.items():
        if rows < 1024 or (kind == wire.FIXED32_F64 and rows <= 65536):
            # below the packing threshold (or few enough distinct values that the encoder would prefer a dictionary):
            # build the packed image by hand
            if kind == wire.FIXED32_F64:
                p = wire.Packed(kind, np.rint(a * 100).astype(np.int32), None, 100.0, "f64")
            else:
                tab, inv = np.unique(a, return_inverse=True)
                if len(tab) == 0:
                    tab = np.zeros(1, dtype=a.dtype)
                ct = np.uint8 if kind in (wire.DICT8_F64, wire.DICT8_I32) else np.uint16
                p = wire.Packed(kind, inv.astype(ct), tab, 0.0, "f64" if a.dtype == np.float64 else "i32")
        else:
            p = wire.pack(a, "f64" if a.dtype == np.float64 else "i32")
            assert p is not None and p.kind == kind, (kind, None if p is None else p.kind)
        ptr, hold, h2d = be.upload_packed(p)
        torch.cuda.synchronize()
        out = hold[:rows * a.dtype.itemsize].cpu().numpy().view(a.dtype)
        assert (bits(out) == bits(a)).all(), wire.KIND_NAMES[kind]
        assert h2d >= p.codes.nbytes


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [0, 1, 5, 1024, 1027, (1 << 20) + 1, 3_000_001])
def test_device_decode_bits_bit_exact(rows):
    """bit-packed kinds through the C ABI (sdqlb200_wire_decode_bits) against the original columns"""
    import torch
    from sdqlpy_b200 import runtime
    runtime.set_backend(None)
    be = runtime.backend()
    rng = np.random.default_rng(rows + 1)
    cases = []
    for nb in (1, 4, 6, 12, 15):
        tab = np.sort(rng.choice(10**6, 2**nb, replace=False)).astype(np.float64) / 100.0
        c = rng.integers(0, 2**nb, rows)
        cases.append((wire.Packed(wire.BITS_DICT_F64, wire.bitpack(c, nb), tab, 0.0, "f64", nbits=nb, rows=rows), tab[c]))
        tabi = (19920101 + np.arange(2**nb)).astype(np.int32)
        cases.append((wire.Packed(wire.BITS_DICT_I32, wire.bitpack(c, nb), tabi, 0.0, "i32", nbits=nb, rows=rows), tabi[c]))
    for nb, base in ((24, 90000), (31, -99999), (32, -2**31), (9, 0)):
        c = rng.integers(0, 2**nb, rows, dtype=np.int64)
        if nb == 32:
            c = np.minimum(c, 2**32 - 1)
        cases.append((wire.Packed(wire.BITS_FIXED_F64, wire.bitpack(c, nb), None, 100.0, "f64", nbits=nb, base=base, rows=rows),
                      (c + base).astype(np.float64) / 100.0))
    for nb, base in ((26, 1), (17, -5), (28, 7)):
        c = rng.integers(0, 2**nb, rows, dtype=np.int64)
        cases.append((wire.Packed(wire.BITS_I32, wire.bitpack(c, nb), None, 0.0, "i32", nbits=nb, base=base, rows=rows),
                      (c + base).astype(np.int32)))
    for nb in (1, 2, 3, 7):
        c = rng.integers(0, 2**nb, rows).astype(np.uint8)
        cases.append((wire.Packed(wire.BITS_U8, wire.bitpack(c, nb), None, 0.0, "code", nbits=nb, rows=rows), c))
    for p, want in cases:
        assert (bits(p.decode_host()) == bits(want)).all()
        ptr, hold, h2d = be.upload_packed(p)
        torch.cuda.synchronize()
        out = hold[:rows * want.dtype.itemsize].cpu().numpy().view(want.dtype)
        assert (bits(out) == bits(want)).all(), (wire.KIND_NAMES[p.kind], p.nbits)


@pytest.mark.gpu
def test_decode_rejects_bad_arguments():
    L = wire.lib()
    assert L.sdqlb200_wire_decode(99, None, None, 0, None, 0.0, None) != 0
    assert L.sdqlb200_wire_decode(wire.DICT8_F64, 16, 32, 4, None, 0.0, None) != 0      # no table
    assert L.sdqlb200_wire_decode(wire.FIXED32_F64, 16, 33, 4, None, 100.0, None) != 0   # misaligned
    assert b"aligned" in L.sdqlb200_wire_last_error()
    assert L.sdqlb200_wire_decode_bits(wire.DICT8_F64, 16, 32, 4, 3, None, 0, 1.0, None) != 0        # not a bit-packed kind
    assert L.sdqlb200_wire_decode_bits(wire.BITS_I32, 16, 32, 4, 33, None, 0, 1.0, None) != 0        # nbits
    assert L.sdqlb200_wire_decode_bits(wire.BITS_DICT_I32, 16, 32, 4, 3, None, 0, 1.0, None) != 0    # no table
    assert L.sdqlb200_wire_decode_bits(wire.BITS_U8, 16, 32, 4, 9, None, 0, 1.0, None) != 0          # > 8 bits of uint8


@pytest.mark.gpu
@pytest.mark.parametrize("q", ["q1", "q6", "q3", "q9", "q18"])
def test_queries_on_packed_columns_gpu(q):
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import runtime
    from util import QUERY_SCRIPT, golden
    runtime.set_backend(None)
    runtime.STORE.clear()
    mod = runtime.load_compiled(QUERY_SCRIPT)
    db = packed_db(0.05, rr.QUERY_ARGS[q])
    assert compare(mod.run(q, db), golden(0.05)[q]) is None
    runtime.STORE.clear()
