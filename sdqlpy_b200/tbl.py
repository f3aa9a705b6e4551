"""`.tbl` reader with device-side parsing: host side of libsdqlb200_tbl.so (include/sdqlb200_tbl.h).

Replaces the reference's ``read_csv`` (sdql_lib.py:118-128 -> 69-115: a Python ``csv.reader`` row loop with per-field
``int()`` / ``float()`` / ``int(v.replace("-", ""))`` conversions into int64 / float64 / ``<U n`` numpy columns; the
TPC-H tables are loaded with it in test_all.py:35-42).  The file's bytes go to the device unchanged; the device indexes
the rows and parses the wanted columns straight into the resident layout (int32 / fp64 / fixed-width bytes).  The
parsed columns come back as compact host ``Column`` objects (string result fields are gathered from them), and the
device copies are handed to the column store, so the first query does not upload them again.

No CPU fallback: without the built library (or without a CUDA device) this raises.
"""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
TBL_SO = os.path.join(PKG, "_build", "libsdqlb200_tbl.so")
T_INT, T_FLOAT, T_DATE, T_STR = 0, 1, 2, 3
MAX_COLS = 32
ERRORS = {1: "the row has fewer fields than the schema", 2: "not a plain decimal number / date",
          3: "integer outside int32 or more than 15 significant digits", 4: "character >= 0x80 in a string field"}


class TblCol(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("width", ctypes.c_int32), ("out", ctypes.c_void_p)]


class TblStatus(ctypes.Structure):
    _fields_ = [("bad_row", ctypes.c_int64), ("error", ctypes.c_int64),
                ("min", ctypes.c_int64 * MAX_COLS), ("max", ctypes.c_int64 * MAX_COLS)]


_lib = None


def lib(path=None):
    """libsdqlb200_tbl.so; raises if it was not built (no CPU fallback on the product path)."""
    global _lib
    if path is not None or _lib is None:
        so = path or TBL_SO
        if not os.path.exists(so):
            raise ImportError("%s not found (run __graft_entry__.build() / sdqlpy_b200.build.compile_tbl())" % so)
        L = ctypes.CDLL(so)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.sdqlb200_tbl_scratch_bytes.argtypes = [i64]
        L.sdqlb200_tbl_scratch_bytes.restype = i64
        L.sdqlb200_tbl_index.argtypes = [vp, i64, vp, ctypes.POINTER(i64), vp]
        L.sdqlb200_tbl_row_starts.argtypes = [vp, i64, vp, vp, i64, vp]
        L.sdqlb200_tbl_parse.argtypes = [vp, vp, i64, ctypes.POINTER(TblCol), ctypes.c_int32, ctypes.c_char, vp, vp]
        L.sdqlb200_tbl_last_error.restype = ctypes.c_char_p
        if path is not None:
            return L
        _lib = L
    return _lib


def _check(L, rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, L.sdqlb200_tbl_last_error().decode()))


def schema_types(schema):
    """[(column, kind)] with kind 'int' | 'float' | 'date' | ('str', n)  ->  [(type id, width)]"""
    out = []
    for name, kind in schema:
        if isinstance(kind, (tuple, list)):
            out.append((T_STR, int(kind[1])))
        elif kind in ("int", "bool"):
            out.append((T_INT, 0))
        elif kind == "float":
            out.append((T_FLOAT, 0))
        elif kind == "date":
            out.append((T_DATE, 0))
        else:
            raise ValueError("column %s: unsupported type %r" % (name, kind))
    return out


def parse_text(text, schema, want=None, delimiter="|", be=None, library=None, row_base=0):
    """Parse one block of `.tbl` text (uint8 numpy array holding whole rows) on the device.
    -> (rows, {column: (kind, host numpy array, width)}, {column: DeviceColumn})"""
    from . import runtime
    be = be or runtime.backend()
    L = library or lib()
    if len(schema) > MAX_COLS:
        raise ValueError("at most %d columns" % MAX_COLS)
    types = schema_types(schema)
    text = np.ascontiguousarray(text, dtype=np.uint8)
    nbytes = int(text.nbytes)
    d_text, h_text = be.upload(text)
    d_scr, h_scr = be.alloc(L.sdqlb200_tbl_scratch_bytes(nbytes))
    rows = ctypes.c_int64(0)
    _check(L, L.sdqlb200_tbl_index(d_text, nbytes, d_scr, ctypes.byref(rows), be.stream()), "sdqlb200_tbl_index")
    n = int(rows.value)
    d_starts, h_starts = be.alloc((n + 1) * 8)
    _check(L, L.sdqlb200_tbl_row_starts(d_text, nbytes, d_scr, d_starts, n, be.stream()), "sdqlb200_tbl_row_starts")
    cols = (TblCol * len(schema))()
    outs = {}
    for i, ((name, _), (t, w)) in enumerate(zip(schema, types)):
        cols[i].type, cols[i].width, cols[i].out = t, w, None
        if want is not None and name not in want:
            continue
        elem = {T_INT: 4, T_DATE: 4, T_FLOAT: 8, T_STR: w}[t]
        ptr, hold = be.alloc(max(1, n) * elem)
        cols[i].out = ptr
        outs[name] = (i, t, w, ptr, hold)
    d_st, h_st = be.alloc(ctypes.sizeof(TblStatus))
    _check(L, L.sdqlb200_tbl_parse(d_text, d_starts, n, cols, len(schema), delimiter.encode("latin1"), d_st, be.stream()),
           "sdqlb200_tbl_parse")
    be.sync()
    st = TblStatus.from_buffer_copy(be.to_host(h_st, ctypes.sizeof(TblStatus)).tobytes())
    if st.bad_row >= 0:
        raise ValueError("malformed .tbl row %d: %s" % (row_base + st.bad_row, ERRORS.get(int(st.error), "error %d" % st.error)))
    host, dev = {}, {}
    for name, (i, t, w, ptr, hold) in outs.items():
        if t == T_STR:
            a = be.to_host(hold, n * w).reshape(n, w).copy()
            host[name] = ("bytes", a, w)
            dev[name] = runtime.DeviceColumn("bytes", ptr, hold, n, 0, 0, w, None, n * w)
        elif t == T_FLOAT:
            a = be.to_host(hold, n * 8).view(np.float64).copy()
            host[name] = ("f64", a, None)
            dev[name] = runtime.DeviceColumn("f64", ptr, hold, n, 0, 0, 8, None, n * 8)
        else:
            a = be.to_host(hold, n * 4).view(np.int32).copy()
            host[name] = ("i32", a, None)
            mn, mx = (int(st.min[i]), int(st.max[i])) if n else (0, 0)
            dev[name] = runtime.DeviceColumn("i32", ptr, hold, n, mn, mx, 4, None, n * 4)
    return n, host, dev


def _no_blank_lines(buf):
    """csv.reader -- the reference's read_csv, sdql_lib.py:79-82 -- yields [] for a blank line and the row loop adds nothing
    for it: blank lines are skipped.  The device index counts every newline as a row end, so they are removed here (one
    memmem over the block; dbgen output has none, so nothing is copied in the normal case)."""
    if buf[:1] == b"\n" or b"\n\n" in buf:
        import re
        buf = re.sub(rb"\n\n+", b"\n", buf).lstrip(b"\n")
    return buf


def read_blocks(path, block_bytes):
    """the file as blocks of whole rows (uint8 arrays), each at most ~block_bytes; blank lines removed"""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        carry = b""
        while True:
            buf = f.read(block_bytes)
            if not buf:
                carry = _no_blank_lines(carry)
                if carry:
                    yield np.frombuffer(carry, dtype=np.uint8)
                return
            buf = carry + buf
            if f.tell() >= size:
                buf = _no_blank_lines(buf)
                if buf:
                    yield np.frombuffer(buf, dtype=np.uint8)
                return
            cut = buf.rfind(b"\n")
            if cut < 0:
                carry = buf
                continue
            head = _no_blank_lines(buf[:cut + 1])
            if head:
                yield np.frombuffer(head, dtype=np.uint8)
            carry = buf[cut + 1:]


def parse_file(path, schema, want=None, delimiter="|", block_bytes=1 << 30, be=None, library=None):
    """-> {column: tpch.gen.Column} (compact host columns) for the wanted columns of a `.tbl` file, parsed on the device.
    A file that fits one block leaves its device-resident columns in the column store (no second upload)."""
    from . import runtime
    from .tpch.gen import Column
    pieces, devs, total = [], [], 0
    for block in read_blocks(path, block_bytes):
        n, host, dev = parse_text(block, schema, want, delimiter, be, library, row_base=total)
        pieces.append(host)
        devs.append(dev)
        total += n
    names = [c for c, _ in schema if want is None or c in want]
    out = {}
    if not pieces:  # empty file
        n, host, dev = parse_text(np.zeros(0, dtype=np.uint8), schema, want, delimiter, be, library)
        pieces, devs = [host], [dev]
    for name in names:
        kind, _, w = pieces[0][name]
        data = pieces[0][name][1] if len(pieces) == 1 else np.concatenate([p[name][1] for p in pieces])
        col = Column(name, kind, data, None, w)
        out[name] = col
        if len(pieces) == 1 and runtime.STORE.enabled:
            d = devs[0][name]
            runtime.STORE.cache[(runtime.STORE.key(col), kind, w if kind == "bytes" else 0)] = (d, col)
    return out


def format_tbl(schema, columns, delimiter="|"):
    """`.tbl` text (bytes) of a table given as reference-layout columns (int64 / float64 / ``<U n`` arrays in schema
    order, e.g. tpch.gen.TPCH.ref_table): dbgen's conventions -- money with two decimals, dates YYYY-MM-DD, a
    trailing delimiter per row.  Used to write test fixtures and to export generated tables."""
    n = len(columns[0]) if columns else 0
    parts = []
    if schema and schema[-1][0].endswith("_NA"):
        # the reference's schemas end with a dummy string(1) column that absorbs the empty field behind dbgen's trailing
        # delimiter (test_all.py:26-33): it is that trailing delimiter, not a field of its own
        schema, columns = schema[:-1], columns[:len(schema) - 1]
    for (name, kind), col in zip(schema, columns):
        if isinstance(kind, (tuple, list)):
            parts.append([str(v) for v in col] if len(col) == n else [""] * n)
        elif kind == "float":
            parts.append(["%.2f" % v for v in col])
        elif kind == "date":
            parts.append(["%04d-%02d-%02d" % (v // 10000, v // 100 % 100, v % 100) for v in col])
        else:
            parts.append([str(int(v)) for v in col])
    lines = [delimiter.join(f) + delimiter + "\n" for f in zip(*parts)]
    return "".join(lines).encode("latin1")
