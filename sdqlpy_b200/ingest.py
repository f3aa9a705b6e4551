"""Reference-layout host columns -> resident device columns, converted ON THE DEVICE (csrc/sdqlb200_ingest.cu, C ABI in
include/sdqlb200_ingest.h).

The reference's generated module borrows the numpy buffers of ``db`` -- int64 for int / date, float64, UCS4 ``<U n`` for
string(n) -- by casting ``PyArray_DATA`` (sdql_compiler.py:644-668).  Here the raw buffer crosses PCIe once, in chunks
through a staging buffer, and is narrowed (int64 -> int32 with min / max), cut to bytes (``<U n`` -> n bytes) or dictionary
encoded (``<U n`` -> uint8 / int32 codes, every row checked against its dictionary entry) by HBM-speed kernels instead of
numpy ``astype`` / ``np.unique`` passes on the host.  For a rank with enough host threads to itself, big int64 / ``<U1``
columns are narrowed by host threads of the same library while the fp64 columns cross the link (``HostNarrow`` below,
``runtime.ColumnStore.get_many``).  PyTorch provides device memory and the copies only."""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(PKG, "_build", "libsdqlb200_ingest.so")
CHUNK_BYTES = int(os.environ.get("SDQLB200_INGEST_CHUNK", str(512 << 20)))  # raw bytes per staged chunk
DICT_SLOTS = 1 << 17    # distinct values of a dictionary-coded string column handled on the device: <= 65536

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError("%s not found (run __graft_entry__.build())" % SO)
        L = ctypes.CDLL(SO)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.sdqlb200_ingest_i64.argtypes = [vp, vp, i64, vp, vp]
        L.sdqlb200_ingest_ucs4_bytes.argtypes = [vp, vp, i64, i32, i32, vp, vp]
        L.sdqlb200_ingest_ucs4_distinct.argtypes = [vp, i64, i64, i32, vp, vp, i64, vp, vp]
        L.sdqlb200_ingest_ucs4_codes.argtypes = [vp, i64, i32, vp, vp, i64, vp, vp, i32, vp, vp]
        L.sdqlb200_ingest_remap.argtypes = [vp, vp, vp, i32, i64, vp]
        L.sdqlb200_ingest_remap_u8.argtypes = [vp, vp, vp, i64, vp]
        L.sdqlb200_ingest_host_i64.argtypes = [vp, vp, i64, i32, vp]
        L.sdqlb200_ingest_host_ucs4_1.argtypes = [vp, vp, i64, i32, vp, vp]
        L.sdqlb200_ingest_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().sdqlb200_ingest_last_error().decode()))


class TooManyValues(Exception):
    """a string column has more distinct values than the device-side dictionary encoder takes"""


def _chunks(a, rows_per):
    for lo in range(0, len(a), rows_per):
        yield lo, a[lo:lo + rows_per]


def _stage(be, t, stage, chunk):
    """host chunk (numpy, any layout) -> the staging tensor's first bytes; asynchronous when the host memory is pinned"""
    src = t.from_numpy(np.ascontiguousarray(chunk).view(np.uint8).reshape(-1))
    dst = stage[:src.numel()]
    dst.copy_(src, non_blocking=src.is_pinned())
    return dst


def upload(a, rep, width, be):
    """numpy column in the reference layout -> (device pointer, holder, min, max, element bytes, dictionary, bytes over the
    link).  ``be``: runtime.CudaBackend."""
    t = be.torch
    L = lib()
    st = be.stream()
    a = np.asarray(a)
    n = len(a)
    if rep == "f64":
        if a.dtype != np.float64:
            a = a.astype(np.float64)
        ptr, holder = be.upload(a)
        return ptr, holder, 0, 0, 8, None, a.nbytes
    if rep == "i32":
        if a.dtype.kind not in "iu":
            raise ValueError("integer column expected, got %s" % a.dtype)
        if a.dtype != np.int64:
            a = a.astype(np.int64)
        out = t.empty(max(n, 4), dtype=t.int32, device=be.dev)
        mm = t.tensor([2**63 - 1, -2**63], dtype=t.int64, device=be.dev)
        per = max(4, (CHUNK_BYTES // 8) & ~3)
        stage = t.empty(min(max(n, 4), per) * 8, dtype=t.uint8, device=be.dev)
        for lo, ch in _chunks(a, per):
            d = _stage(be, t, stage, ch)
            _ck(L.sdqlb200_ingest_i64(d.data_ptr(), out.data_ptr() + 4 * lo, len(ch), mm.data_ptr(), st), "ingest_i64")
        mn, mx = (int(x) for x in mm.cpu()) if n else (0, 0)
        if n and (mn < -2**31 or mx >= 2**31):
            raise ValueError("integer column outside int32 range (device layout is int32 in this version)")
        return out.data_ptr(), out, mn, mx, 4, None, a.nbytes
    if a.dtype.kind != "U":
        raise ValueError("string column expected, got %s" % a.dtype)
    nchar = a.dtype.itemsize // 4
    raw = np.ascontiguousarray(a).view(np.uint32).reshape(n, nchar)
    per = max(1, CHUNK_BYTES // (4 * nchar))
    stage = t.empty(min(max(n, 1), per) * 4 * nchar, dtype=t.uint8, device=be.dev)
    bad = t.full((1,), 2**63 - 1, dtype=t.int64, device=be.dev)
    if rep == "bytes":
        out = t.zeros((max(n, 1), width), dtype=t.uint8, device=be.dev)
        for lo, ch in _chunks(raw, per):
            d = _stage(be, t, stage, ch)
            _ck(L.sdqlb200_ingest_ucs4_bytes(d.data_ptr(), out.data_ptr() + lo * width, len(ch), nchar, width, bad.data_ptr(), st),
                "ingest_ucs4_bytes")
        if n and int(bad.cpu()) != 2**63 - 1:
            raise ValueError("non-latin1 characters are not supported in string columns (row %d)" % (int(bad.cpu()) - 1))
        return out.data_ptr(), out, 0, 0, width, None, raw.nbytes
    if rep != "code":
        raise ValueError(rep)
    # dictionary encoding, chunk by chunk: distinct set (device) -> new values get the next provisional code (host, a few
    # strings) -> codes of the chunk's rows (device, each row compared with its entry) -> final order at the end
    cap = DICT_SLOTS
    keys = t.full((cap,), -1, dtype=t.int64, device=be.dev)
    reps = t.full((cap,), 2**63 - 1, dtype=t.int64, device=be.dev)
    count = t.zeros(1, dtype=t.int64, device=be.dev)
    slot_code_h = np.full(cap, -1, dtype=np.int32)
    values = []
    prov = t.empty(max(n, 1), dtype=t.int32, device=be.dev)
    for lo, ch in _chunks(raw, per):
        d = _stage(be, t, stage, ch)
        _ck(L.sdqlb200_ingest_ucs4_distinct(d.data_ptr(), len(ch), lo, nchar, keys.data_ptr(), reps.data_ptr(), cap,
                                            count.data_ptr(), st), "ingest_ucs4_distinct")
        if int(count.cpu()) >= cap // 2:
            raise TooManyValues()
        if int(count.cpu()) > len(values):
            kh, rh = keys.cpu().numpy(), reps.cpu().numpy()
            new = np.nonzero((kh != -1) & (slot_code_h < 0))[0]
            new = new[np.argsort(rh[new], kind="stable")]  # first-seen order: deterministic provisional codes
            for s_ in new:
                slot_code_h[s_] = len(values)
                values.append(raw[rh[s_]].copy())
        sc = t.from_numpy(slot_code_h).to(be.dev)
        dd = t.from_numpy(np.ascontiguousarray(np.stack(values)).view(np.int32)).to(be.dev)
        _ck(L.sdqlb200_ingest_ucs4_codes(d.data_ptr(), len(ch), nchar, keys.data_ptr(), sc.data_ptr(), cap, dd.data_ptr(),
                                         prov.data_ptr() + 4 * lo, 4, bad.data_ptr(), st), "ingest_ucs4_codes")
        if int(bad.cpu()) != 2**63 - 1:
            raise RuntimeError("dictionary encoding: row %d does not equal its dictionary entry (hash collision)" % (lo + int(bad.cpu()) - 1))
    strs = [str(np.ascontiguousarray(v).view("<U%d" % nchar)[0]) for v in values]
    order = sorted(range(len(strs)), key=lambda i: strs[i])  # np.unique order: code points ascending
    table = np.zeros(max(1, len(strs)), dtype=np.int32)
    for newc, i in enumerate(order):
        table[i] = newc
    dictionary = [strs[i] for i in order]
    w = 1 if len(dictionary) <= 256 else 4
    out = t.empty(max(n, 4), dtype=t.uint8 if w == 1 else t.int32, device=be.dev)
    tb = t.from_numpy(table).to(be.dev)
    _ck(L.sdqlb200_ingest_remap(prov.data_ptr(), tb.data_ptr(), out.data_ptr(), w, n, st), "ingest_remap")
    t.cuda.current_stream().synchronize()
    return out.data_ptr(), out, 0, max(0, len(dictionary) - 1), w, dictionary, raw.nbytes


def recode(col_holder, n, width, table, be):
    """replace the codes of a resident dictionary-coded column by table[code] (a rank's dictionary -> the dictionary all
    ranks agreed on); -> (pointer, holder, element bytes)"""
    t = be.torch
    src = col_holder[:n].to(t.int32)
    w = 1 if len(table) and int(table.max()) < 256 else (4 if len(table) else width)  # the merged dictionary's code width
    out = t.empty(max(n, 4), dtype=t.uint8 if w == 1 else t.int32, device=be.dev)
    tb = t.from_numpy(np.ascontiguousarray(table, dtype=np.int32)).to(be.dev)
    _ck(lib().sdqlb200_ingest_remap(src.data_ptr(), tb.data_ptr(), out.data_ptr(), w, n, be.stream()), "ingest_remap")
    return out.data_ptr(), out, w


# ---------------------------------------------------------------------------------------------
# host-side narrowing in front of the upload (include/sdqlb200_ingest.h, "HOST side of the upload")
# ---------------------------------------------------------------------------------------------
_staging = {}  # (device, dtype, rows) -> page-locked staging tensors (reused: pinning gigabytes is slow)


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    return max(1, min(n, int(os.environ.get("SDQLB200_HOST_THREADS", "16"))))


class HostNarrow:
    """int64 -> int32 / `<U1` -> byte narrowing of one host column by host threads, started in the background (the C call
    releases the GIL) so that columns which need no host work cross the link meanwhile; finish() uploads the narrowed image.
    -> (device pointer, holder, min, max, element bytes, dictionary, bytes over the link), or None when the column turned out
    not to be narrowable (a `<U1` code point > 255: the device-side encoder takes it)."""

    def __init__(self, a, rep, be):
        import threading
        t = be.torch
        self.a, self.rep, self.be, self.n = np.ascontiguousarray(a), rep, be, len(a)
        dt = t.int32 if rep == "i32" else t.uint8
        key = (str(be.dev), str(dt), self.n)
        pool = _staging.setdefault(key, [])
        self.stage = pool.pop() if pool else t.empty(max(self.n, 4), dtype=dt, pin_memory=True)
        self.key = key
        self.mm = np.zeros(2, dtype=np.int64)
        self.present = np.zeros(4, dtype=np.uint64)
        self.bad = np.full(1, -1, dtype=np.int64)
        self.rc = None
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        L = lib()
        if self.rep == "i32":
            self.rc = L.sdqlb200_ingest_host_i64(self.a.ctypes.data, self.stage.data_ptr(), self.n, host_threads(), self.mm.ctypes.data)
        else:
            self.rc = L.sdqlb200_ingest_host_ucs4_1(self.a.ctypes.data, self.stage.data_ptr(), self.n, host_threads(),
                                                    self.present.ctypes.data, self.bad.ctypes.data)

    def finish(self):
        self.th.join()
        t, be, n = self.be.torch, self.be, self.n
        try:
            _ck(self.rc, "ingest_host")
            if self.rep == "i32":
                mn, mx = (int(self.mm[0]), int(self.mm[1])) if n else (0, 0)
                if n and (mn < -2**31 or mx >= 2**31):
                    raise ValueError("integer column outside int32 range (device layout is int32 in this version)")
                out = t.empty(max(n, 4), dtype=t.int32, device=be.dev)
                out[:n].copy_(self.stage[:n], non_blocking=True)
                self._release(out)
                return out.data_ptr(), out, mn, mx, 4, None, 4 * n
            if int(self.bad[0]) >= 0:
                self._release(None)
                return None
            values = [b for b in range(256) if (int(self.present[b >> 6]) >> (b & 63)) & 1]
            dictionary = ["" if b == 0 else chr(b) for b in values]  # np.unique order: code points ascending
            table = np.zeros(256, dtype=np.uint8)
            for code, b in enumerate(values):
                table[b] = code
            out = t.empty(max(n, 4), dtype=t.uint8, device=be.dev)
            out[:n].copy_(self.stage[:n], non_blocking=True)
            tb = t.from_numpy(table).to(be.dev)
            _ck(lib().sdqlb200_ingest_remap_u8(out.data_ptr(), tb.data_ptr(), out.data_ptr(), n, be.stream()), "ingest_remap_u8")
            self._release(out)
            return out.data_ptr(), out, 0, max(0, len(dictionary) - 1), 1, dictionary, n
        except Exception:
            self._release(None)
            raise

    def _release(self, dev_tensor):
        """the staging buffer goes back to the pool once the copy out of it has run"""
        if dev_tensor is not None:
            self.be.torch.cuda.current_stream().synchronize()
        _staging.setdefault(self.key, []).append(self.stage)
        self.stage = None
