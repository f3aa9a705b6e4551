#!/usr/bin/env python3
"""Throughput of the device-side `.tbl` reader (csrc/sdqlb200_tbl.cu): a lineitem text image of ~--mb megabytes (the rows of
a small generated table repeated -- parsing cost does not depend on the values) is parsed on cuda:0; CUDA events around the
three device steps (row index, row starts, field parsing into the resident columns) with the text already in HBM, and the
wall time of the whole parse_text call from host memory (upload over PCIe + parse + the columns back to the host).

   python tools/bench_tbl.py --mb 2048 [--out profiles/...json]
The reference's read_csv (sdql_lib.py:69-128) is a Python csv.reader loop: timed on a 1/64 sample of the same text with
the reference package when oracle/_ref/site holds it (TEST INFRASTRUCTURE).
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    from sdqlpy_b200 import runtime, tbl
    from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH
    g = TPCH(0.01)
    schema = SCHEMAS["lineitem"]
    block = tbl.format_tbl(schema, g.ref_table("lineitem", [c for c, _ in schema]))
    reps = max(1, (a.mb << 20) // len(block))
    text = np.frombuffer(block * reps, dtype=np.uint8)
    nbytes = int(text.nbytes)
    be = runtime.backend()
    L = tbl.lib()
    types = tbl.schema_types(schema)
    want = [c for c, k in schema if not c.endswith("_NA") and c != "l_comment"]
    d_text, h_text = be.upload(text)
    d_scr, h_scr = be.alloc(L.sdqlb200_tbl_scratch_bytes(nbytes))
    rows = ctypes.c_int64(0)
    st = be.stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    best = None
    for _ in range(a.reps):
        torch.cuda.synchronize()
        ev[0].record()
        tbl._check(L, L.sdqlb200_tbl_index(d_text, nbytes, d_scr, ctypes.byref(rows), st), "index")
        ev[1].record()
        n = int(rows.value)
        d_starts, h_starts = be.alloc((n + 1) * 8)
        tbl._check(L, L.sdqlb200_tbl_row_starts(d_text, nbytes, d_scr, d_starts, n, st), "row_starts")
        ev[2].record()
        cols = (tbl.TblCol * len(schema))()
        keep, out_bytes = [], 0
        for i, ((name, _), (t, w)) in enumerate(zip(schema, types)):
            cols[i].type, cols[i].width, cols[i].out = t, w, None
            if name in want:
                elem = {tbl.T_INT: 4, tbl.T_DATE: 4, tbl.T_FLOAT: 8, tbl.T_STR: w}[t]
                ptr, hold = be.alloc(max(1, n) * elem)
                cols[i].out = ptr
                keep.append(hold)
                out_bytes += n * elem
        d_st, h_st = be.alloc(ctypes.sizeof(tbl.TblStatus))
        torch.cuda.synchronize()
        ev[2].record()
        tbl._check(L, L.sdqlb200_tbl_parse(d_text, d_starts, n, cols, len(schema), b"|", d_st, st), "parse")
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
        if best is None or sum(t) < sum(best):
            best = t
        del keep
    dev_ms = sum(best)
    # the whole call from host memory
    t0 = time.perf_counter()
    n2, host, dev = tbl.parse_text(text, schema, want)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    row = {"text_bytes": nbytes, "rows": n, "columns_parsed": len(want), "resident_bytes_written": out_bytes,
           "index_ms": round(best[0], 3), "row_starts_ms": round(best[1], 3), "parse_ms": round(best[2], 3),
           "device_ms": round(dev_ms, 3), "device_text_GBps": round(nbytes / dev_ms / 1e6, 1),
           "device_traffic_GBps": round((3 * nbytes + out_bytes) / dev_ms / 1e6, 1),
           "traffic_model": "text read by the three steps + resident columns written",
           "call_from_host_s": round(wall, 3), "call_text_GBps": round(nbytes / wall / 1e9, 2),
           "note": "call_from_host: text over PCIe (pageable numpy buffer) + device parse + parsed columns copied back to the host"}
    try:  # the reference's reader on a sample (Python loop: ~1 MB/s per core)
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "site"))
        import tempfile
        from sdqlpy.sdql_lib import date, read_csv, string  # noqa: E402 -- the real reference package (test infrastructure)
        _b, _f, _i = bool, float, int  # the reference's schemas use the builtins for these (test_all.py:26-33)
        sample = block * max(1, reps // 64)
        with tempfile.NamedTemporaryFile(suffix=".tbl", delete=False) as f:
            f.write(sample)
            path = f.name
        ty = {}
        for c, k in schema:
            ty[c] = string(k[1]) if isinstance(k, tuple) else {"int": _i, "float": _f, "date": date}[k]
        t0 = time.perf_counter()
        read_csv(path, {record_key(ty): _b}, "li")
        dt = time.perf_counter() - t0
        os.unlink(path)
        row["reference_read_csv"] = {"sample_bytes": len(sample), "seconds": round(dt, 2), "GBps": round(len(sample) / dt / 1e9, 4)}
    except Exception as ex:  # noqa: BLE001
        row["reference_read_csv"] = "unavailable: %r" % (ex,)
    print(json.dumps(row))
    if a.out:
        json.dump(row, open(a.out, "w"), indent=1)


def record_key(ty):
    from sdqlpy.sdql_lib import record
    return record(ty)


if __name__ == "__main__":
    main()
