"""Device-side TPC-H fact tables: lineitem / orders generated straight into HBM (csrc/sdqlb200_tpchgen.cu, C ABI in
include/sdqlb200_tpchgen.h), bit-identical to the host generator ``gen.TPCH`` and directly usable as query inputs
(``runtime.DeviceColumn``).  Stands where the reference has ``read_csv`` of dbgen files (sdql_lib.py:69-128): at
SF100 lineitem is 600 M rows -- it never exists on the host.  The small dimension tables stay host generated.

PyTorch provides device memory and the prefix sum only."""
import ctypes
import os

import numpy as np

from .. import runtime
from . import gen

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN_SO = os.path.join(PKG, "_build", "libsdqlb200_tpchgen.so")

LI_I32 = ["l_orderkey", "l_partkey", "l_suppkey", "l_linenumber", "l_shipdate", "l_commitdate", "l_receiptdate"]
LI_F64 = ["l_quantity", "l_extendedprice", "l_discount", "l_tax"]
LI_CODE = {"l_returnflag": (gen.RFLAGS, 1), "l_linestatus": (gen.LSTATUS, 1), "l_shipinstruct": (gen.INSTRUCTS, 25),
           "l_shipmode": (gen.MODES, 10)}
ORD_I32 = ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"]
ORD_F64 = ["o_totalprice"]
ORD_CODE = {"o_orderstatus": (gen.OSTATUS, 1), "o_orderpriority": (gen.PRIORITIES, 15)}


class Params(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_int64), ("S", ctypes.c_int64), ("P", ctypes.c_int64), ("C", ctypes.c_int64),
                ("O", ctypes.c_int64)]


class LineitemCols(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in LI_I32 + LI_F64 + list(LI_CODE)]


class OrdersCols(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ORD_I32 + ORD_F64 + list(ORD_CODE) + ["o_comment"]]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(GEN_SO):
            raise ImportError("%s not found (run __graft_entry__.build())" % GEN_SO)
        L = ctypes.CDLL(GEN_SO)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.sdqlb200_tpchgen_order_lines.argtypes = [ctypes.POINTER(Params), i64, i64, vp, vp]
        L.sdqlb200_tpchgen_lineitem.argtypes = [ctypes.POINTER(Params), i64, i64, vp, ctypes.POINTER(LineitemCols), vp]
        L.sdqlb200_tpchgen_orders.argtypes = [ctypes.POINTER(Params), i64, i64, vp, ctypes.POINTER(OrdersCols), vp]
        L.sdqlb200_tpchgen_orders_text.argtypes = [ctypes.POINTER(Params), i64, i64, vp, ctypes.POINTER(OrdersCols), vp,
                                                   ctypes.c_int32, vp]
        L.sdqlb200_tpchgen_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


class DeviceTPCH:
    """``DeviceTPCH(sf).columns("lineitem", cols, order_range)`` -> {name: runtime.DeviceColumn}; same sizes, seed and
    values as ``gen.TPCH(sf)``."""

    def __init__(self, sf=1.0, seed=gen.SEED):
        import torch
        self.torch = torch
        self.host = gen.TPCH(sf, seed)
        h = self.host
        self.params = Params(h.seed, h.S, h.P, h.C, h.O)
        self.O = h.O
        self.be = runtime.backend()
        self._off = None
        self._vocab = None

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().sdqlb200_tpchgen_last_error().decode()))

    def offsets(self):
        """int64 device tensor [O + 1]: global lineitem row id of the first line of every order."""
        if self._off is None:
            t = self.torch
            nl = t.empty(self.O, dtype=t.int32, device=self.be.dev)
            self._ck(lib().sdqlb200_tpchgen_order_lines(ctypes.byref(self.params), 0, self.O, nl.data_ptr(), self.be.stream()),
                     "tpchgen_order_lines")
            off = t.zeros(self.O + 1, dtype=t.int64, device=self.be.dev)
            t.cumsum(nl, 0, out=off[1:])
            self._off = off
        return self._off

    def rows(self, table, order_range=None):
        o0, o1 = order_range or (0, self.O)
        if table == "orders":
            return o1 - o0
        off = self.offsets()
        return int(off[o1] - off[o0])

    def _col(self, kind, tensor, rows, dictionary=None, width=None):
        mn = mx = stride = 0
        if kind == "i32" and rows:
            mn, mx = int(tensor.min()), int(tensor.max())
            stride = runtime.stride_stat(tensor[:rows], mn)
        elif kind == "code":
            mn, mx = 0, len(dictionary) - 1
        w = {"i32": 4, "f64": 8, "code": 1}.get(kind, width)
        return runtime.DeviceColumn(kind, tensor.data_ptr(), tensor, rows, mn, mx, w, dictionary, tensor.numel() * tensor.element_size(), stride)

    def columns(self, table, cols=None, order_range=None):
        t = self.torch
        o0, o1 = order_range or (0, self.O)
        off = self.offsets()[o0:o1 + 1].contiguous()
        dev = self.be.dev
        out = {}
        if table == "lineitem":
            n = int(off[-1] - off[0])
            want = [c for c in LI_I32 + LI_F64 + list(LI_CODE) if cols is None or c in cols]
            st, hold = LineitemCols(), {}
            for c in want:
                dt = t.int32 if c in LI_I32 else t.float64 if c in LI_F64 else t.uint8
                hold[c] = t.empty(max(n, 4), dtype=dt, device=dev)
                setattr(st, c, hold[c].data_ptr())
            self._ck(lib().sdqlb200_tpchgen_lineitem(ctypes.byref(self.params), o0, o1, off.data_ptr(), ctypes.byref(st),
                                                     self.be.stream()), "tpchgen_lineitem")
            for c in want:
                if c in LI_CODE:
                    out[c] = self._col("code", hold[c][:n], n, list(LI_CODE[c][0]))
                else:
                    out[c] = self._col("i32" if c in LI_I32 else "f64", hold[c][:n], n)
            return out
        if table == "orders":
            n = o1 - o0
            want = [c for c in ORD_I32 + ORD_F64 + list(ORD_CODE) + ["o_comment"] if cols is None or c in cols]
            st, hold = OrdersCols(), {}
            for c in want:
                if c == "o_comment":
                    hold[c] = t.empty((max(n, 1), 79), dtype=t.uint8, device=dev)
                else:
                    dt = t.int32 if c in ORD_I32 else t.float64 if c in ORD_F64 else t.uint8
                    hold[c] = t.empty(max(n, 4), dtype=dt, device=dev)
                setattr(st, c, hold[c].data_ptr())
            vocab_ptr, nwords = None, 0
            if "o_comment" in want:
                if self._vocab is None:
                    v = gen._bytes_table([w + " " for w in gen.WORDS], 12)
                    v[v == 0] = 32
                    self._vocab = t.from_numpy(np.ascontiguousarray(v)).to(dev)
                vocab_ptr, nwords = self._vocab.data_ptr(), len(gen.WORDS)
            self._ck(lib().sdqlb200_tpchgen_orders_text(ctypes.byref(self.params), o0, o1, off.data_ptr(), ctypes.byref(st),
                                                        vocab_ptr, nwords, self.be.stream()), "tpchgen_orders")
            for c in want:
                if c == "o_comment":
                    out[c] = runtime.DeviceColumn("bytes", hold[c].data_ptr(), hold[c], n, 0, 0, 79, None, n * 79)
                elif c in ORD_CODE:
                    out[c] = self._col("code", hold[c][:n], n, list(ORD_CODE[c][0]))
                else:
                    out[c] = self._col("i32" if c in ORD_I32 else "f64", hold[c][:n], n)
            return out
        raise ValueError("only lineitem and orders are generated on the device (use gen.TPCH for %s)" % table)
