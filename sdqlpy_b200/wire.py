"""Column wire formats: lossless packed host images of columns for the host -> device link, expanded on the device by
csrc/sdqlb200_wire.cu (C ABI: include/sdqlb200_wire.h).

The reference keeps every column as int64 / float64 / UCS4 in host memory (read_csv, sdql_lib.py:83-97) and the
generated module reads those buffers in place (sdql_compiler.py:644-668).  A B200 reads its columns from HBM, so
they have to cross PCIe once per upload; packing is done once per column at load time (the moment the reference
spends in read_csv), never per query.  Every packed form is verified bit for bit against the column it encodes --
when no packed form is exact the column travels in its resident form.

    dict8 / dict16   value -> code into a sorted dictionary of the distinct values (quantities, discounts, taxes,
                     dates, small integer domains)
    fixed32          decimal(.,2) money: int32 hundredths, decoded as (double)v / 100.0 -- IEEE division gives exactly
                     the double a decimal parser produces for the same digits
    bits_*           the same codes / hundredths / integers as fields of exactly ceil(log2(range)) bits in a little-endian
                     bit stream (frame of reference: value - min); dictionary-coded string columns travel the same way
                     (2 bits for l_returnflag, 1 for l_linestatus).  TPC-H Q1's seven columns: 53 bits per row.
"""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
WIRE_SO = os.path.join(PKG, "_build", "libsdqlb200_wire.so")

DICT8_F64, DICT16_F64, DICT8_I32, DICT16_I32, FIXED32_F64, BITS_DICT_F64, BITS_DICT_I32, BITS_FIXED_F64, BITS_I32, BITS_U8 = range(10)
KIND_NAMES = ["dict8_f64", "dict16_f64", "dict8_i32", "dict16_i32", "fixed32_f64",
              "bits_dict_f64", "bits_dict_i32", "bits_fixed_f64", "bits_i32", "bits_u8"]
SRC_WIDTH = [1, 2, 1, 2, 4, 0, 0, 0, 0, 0]
DST_DTYPE = [np.float64, np.float64, np.int32, np.int32, np.float64, np.float64, np.int32, np.float64, np.int32, np.uint8]
SAMPLE = 1 << 20
BITPACK = os.environ.get("SDQLB200_WIRE_BITS", "1") != "0"  # 0: byte-aligned kinds only
_CHUNK = 1 << 20  # rows per bit-packing step (a multiple of 8: chunks start on byte boundaries)


def bitpack(codes, nbits):
    """non-negative integers < 2**nbits -> little-endian bit stream (uint8), zero padded to 16 bytes + 16 bytes slack
    (the device decoder reads a 64-bit window per field).  8 fields = nbits whole bytes: every group of 8 values is
    assembled in ceil(nbits / 8) 64-bit words with vector shifts."""
    n = len(codes)
    nbytes = (n * nbits + 7) // 8
    out = np.zeros(((nbytes + 15) // 16) * 16 + 16, dtype=np.uint8)
    nw = (8 * nbits + 63) // 64
    for s in range(0, n, _CHUNK):
        c = np.asarray(codes[s:s + _CHUNK]).astype(np.uint64)
        m = len(c)
        if m % 8:
            c = np.concatenate([c, np.zeros(8 - m % 8, dtype=np.uint64)])
        c = c.reshape(-1, 8)
        words = np.zeros((c.shape[0], nw), dtype=np.uint64)
        for j in range(8):
            lo = j * nbits
            w, sh = lo // 64, lo % 64
            words[:, w] |= c[:, j] << np.uint64(sh)  # bits shifted past 64 are dropped ...
            if sh + nbits > 64:                     # ... and land in the next word
                words[:, w + 1] |= c[:, j] >> np.uint64(64 - sh)
        pb = words.view(np.uint8).reshape(c.shape[0], nw * 8)[:, :nbits].reshape(-1)[:(m * nbits + 7) // 8]
        o = s * nbits // 8
        out[o:o + len(pb)] = pb
    return out


def bitunpack(stream, nbits, n):
    """numpy restatement of the device field extraction (tests / emulation side only) -> uint64 codes"""
    out = np.empty(n, dtype=np.uint64)
    w = (np.uint64(1) << np.arange(nbits, dtype=np.uint64))
    for s in range(0, n, _CHUNK):
        m = min(_CHUNK, n - s)
        o = s * nbits // 8
        b = np.unpackbits(stream[o:o + (m * nbits + 7) // 8], bitorder="little")[:m * nbits].reshape(m, nbits)
        out[s:s + m] = (b.astype(np.uint64) * w).sum(axis=1)
    return out


class Packed:
    """one packed column: ``codes`` (the image that crosses the link), ``table`` (dictionary, tiny) or ``scale``."""
    __slots__ = ("kind", "codes", "table", "scale", "rows", "rep", "min", "max", "nbits", "base", "dictionary",
                 "_dev_table", "_pin", "_stage")

    def __init__(self, kind, codes, table=None, scale=0.0, rep="f64", mn=0, mx=0, nbits=0, base=0, rows=None,
                 dictionary=None):
        self.kind, self.codes, self.table, self.scale, self.rep = kind, codes, table, float(scale), rep
        self.rows = len(codes) if rows is None else int(rows)  # bit-packed kinds: codes is the byte stream
        self.min, self.max, self.nbits, self.base, self.dictionary = mn, mx, int(nbits), int(base), dictionary
        self._dev_table, self._pin, self._stage = None, None, None

    @property
    def nbytes(self):
        return self.codes.nbytes + (self.table.nbytes if self.table is not None else 0)

    def pin(self, be):
        """move the packed image to page-locked host memory (asynchronous uploads)."""
        if self._pin is None:
            raw, self._pin = be.pinned_like(self.codes.view(np.uint8))
            self.codes = raw.view(self.codes.dtype)
        return self

    def decode_host(self):
        """numpy restatement of the device kernels (tests / oracle side only)."""
        if self.kind == FIXED32_F64:
            return self.codes.astype(np.float64) / self.scale
        if self.kind < BITS_DICT_F64:
            return self.table[self.codes]
        c = bitunpack(self.codes, self.nbits, self.rows)
        if self.kind in (BITS_DICT_F64, BITS_DICT_I32):
            return self.table[c.astype(np.int64)]
        if self.kind == BITS_FIXED_F64:
            return (c.astype(np.int64) + self.base).astype(np.float64) / self.scale
        if self.kind == BITS_I32:
            return (c.astype(np.int64) + self.base).astype(np.int32)
        return c.astype(np.uint8)


def _bits(a):
    return a.view(np.int64) if a.dtype == np.float64 else a


def _try_dict(a, limit=65536):
    """-> (codes, table) with table[codes] == a bit for bit, or None."""
    n = len(a)
    if n == 0:
        return None
    tab = np.unique(a[:SAMPLE])
    for _ in range(2):
        if len(tab) > limit:
            return None
        codes = np.searchsorted(tab, a)
        np.minimum(codes, len(tab) - 1, out=codes)
        bad = _bits(tab[codes]) != _bits(a)
        if not bad.any():
            ct = np.uint8 if len(tab) <= 256 else np.uint16
            return codes.astype(ct), tab
        extra = np.unique(a[bad])
        if len(extra) + len(tab) > limit or (a.dtype == np.float64 and np.isnan(extra).any()):
            return None
        tab = np.unique(np.concatenate([tab, extra]))
        # -0.0 / +0.0 collapse in np.unique: the bitwise check above rejects such a column on the second pass
    return None


def _try_fixed32(a, scale=100.0):
    if a.dtype != np.float64 or len(a) == 0:
        return None
    with np.errstate(invalid="ignore", over="ignore"):
        v = np.rint(a * scale)
        if not np.isfinite(v).all() or np.abs(v).max() >= 2**31:
            return None
        c = v.astype(np.int32)
        if (_bits(c.astype(np.float64) / scale) != _bits(a)).any():
            return None
    return c


def pack(arr, rep):
    """host column in its resident form (int32 / float64 numpy array) -> Packed, or None when it should travel as is."""
    a = np.ascontiguousarray(arr)
    want = {"i32": np.int32, "f64": np.float64}.get(rep)
    if want is None or a.dtype != want or len(a) < 1024:
        return None
    mn = mx = 0
    if rep == "i32":
        mn, mx = int(a.min()), int(a.max())
    d = _try_dict(a)
    if d is not None:
        codes, tab = d
        wide = codes.dtype == np.uint16
        nb = max(1, int(len(tab) - 1).bit_length())
        if BITPACK and nb not in (8, 16):  # exactly as many bits as the dictionary needs
            return Packed(BITS_DICT_F64 if rep == "f64" else BITS_DICT_I32, bitpack(codes, nb), np.ascontiguousarray(tab),
                          0.0, rep, mn, mx, nbits=nb, rows=len(a))
        kind = (DICT16_F64 if wide else DICT8_F64) if rep == "f64" else (DICT16_I32 if wide else DICT8_I32)
        return Packed(kind, codes, np.ascontiguousarray(tab), 0.0, rep, mn, mx)
    if rep == "f64":
        c = _try_fixed32(a)
        if c is not None:
            lo = int(c.min())
            nb = max(1, (int(c.max()) - lo).bit_length())
            if BITPACK and nb < 32:  # frame of reference: hundredths above the column minimum
                p = Packed(BITS_FIXED_F64, bitpack(c.astype(np.int64) - lo, nb), None, 100.0, rep, nbits=nb, base=lo, rows=len(a))
                if (_bits(((c.astype(np.int64) - lo) + lo).astype(np.float64) / 100.0) == _bits(a)).all():
                    return p
            return Packed(FIXED32_F64, c, None, 100.0, rep)
    elif BITPACK:
        nb = max(1, (mx - mn).bit_length())
        if nb <= 28:  # frame-of-reference integers (keys, sizes): saves at least half a byte per value
            return Packed(BITS_I32, bitpack(a.astype(np.int64) - mn, nb), None, 0.0, rep, mn, mx, nbits=nb, base=mn, rows=len(a))
    return None


def pack_codes(codes, dictionary):
    """dictionary-coded string column (uint8 codes) -> bit-packed image, or None"""
    a = np.ascontiguousarray(codes)
    if not BITPACK or a.dtype != np.uint8 or len(a) < 1024:
        return None
    nb = max(1, int(len(dictionary) - 1).bit_length())
    if nb >= 8:
        return None
    return Packed(BITS_U8, bitpack(a, nb), None, 0.0, "code", 0, len(dictionary) - 1, nbits=nb, rows=len(a),
                  dictionary=list(dictionary))


def pack_column(col):
    """attach the packed image to a tpch.gen.Column (no-op for kinds that are already narrow: codes, bytes)."""
    if getattr(col, "wire", None) is None and col.kind in ("i32", "f64"):
        col.wire = pack(col.data, col.kind)
    elif getattr(col, "wire", None) is None and col.kind == "code" and col.dictionary is not None and len(col.dictionary) <= 256:
        col.wire = pack_codes(col.data, col.dictionary)
    return col


def pack_db(db):
    """pack every column of a db (list of per-relation column lists) once, at load time."""
    for rel in db:
        for c in rel:
            if c is not None and hasattr(c, "kind") and hasattr(c, "wire"):
                pack_column(c)
    return db


# ---------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------
_lib = None


def lib():
    """libsdqlb200_wire.so; raises if it was not built (no CPU fallback on the product path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(WIRE_SO):
            raise ImportError("%s not found (run __graft_entry__.build() / sdqlpy_b200.build.compile_wire())" % WIRE_SO)
        L = ctypes.CDLL(WIRE_SO)
        L.sdqlb200_wire_decode.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
        L.sdqlb200_wire_decode.restype = ctypes.c_int
        L.sdqlb200_wire_decode_bits.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                                ctypes.c_void_p]
        L.sdqlb200_wire_decode_bits.restype = ctypes.c_int
        L.sdqlb200_wire_last_error.restype = ctypes.c_char_p
        L.sdqlb200_wire_src_width.argtypes = [ctypes.c_int32]
        L.sdqlb200_wire_dst_width.argtypes = [ctypes.c_int32]
        _lib = L
    return _lib


def upload_decoded(p, be):
    """Packed -> (device pointer, holder, bytes that crossed the link): upload the packed image on the current stream
    and expand it there."""
    L = lib()
    h2d = p.codes.nbytes
    if p._pin is not None and hasattr(be, "upload_overlapped") and p.codes.nbytes >= (1 << 20):
        # pinned image: the copy runs on the back end's copy stream into a staging buffer that belongs to this column,
        # so the link keeps transferring the next column while this one is expanded on the compute stream
        src_ptr = be.upload_overlapped(p)
    else:
        src_ptr, src_hold = be.upload(p.codes.view(np.uint8))
    tab_ptr = None
    if p.table is not None:
        if p._dev_table is None or p._dev_table[2] is not be:
            tp, th = be.upload(p.table)
            p._dev_table = (tp, th, be)
            h2d += p.table.nbytes
        tab_ptr = p._dev_table[0]
    dst_ptr, dst_hold = be.alloc(p.rows * np.dtype(DST_DTYPE[p.kind]).itemsize)
    if p.kind >= BITS_DICT_F64:
        rc = L.sdqlb200_wire_decode_bits(p.kind, src_ptr, dst_ptr, p.rows, p.nbits, tab_ptr, p.base, p.scale, be.stream())
    else:
        rc = L.sdqlb200_wire_decode(p.kind, src_ptr, dst_ptr, p.rows, tab_ptr, p.scale, be.stream())
    if rc != 0:
        raise RuntimeError("sdqlb200_wire_decode failed (%d): %s" % (rc, L.sdqlb200_wire_last_error().decode()))
    if hasattr(be, "decoded"):
        be.decoded(p)
    # src_hold is dropped here: the caching allocator reuses it in stream order, after the decode kernel
    return dst_ptr, dst_hold, h2d
