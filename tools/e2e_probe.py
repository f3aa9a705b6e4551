#!/usr/bin/env python3
"""Where does the end-to-end step (plain reference-layout numpy columns -> PCIe -> device conversion -> query) spend its
time?  Per column of Q1's input: seconds of the store's upload path (synchronised), against a bare pinned copy of the same
bytes (the PCIe floor on this box).

   python tools/e2e_probe.py --sf 10 [--out profiles/...json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--query", default="q1")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    from bench import host_reference_columns, needed, reference_db
    from sdqlpy_b200 import runtime
    from sdqlpy_b200.tpch.gen_device import DeviceTPCH
    mod = runtime.load_compiled(os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py"))
    man = mod.queries[a.query]
    dg = DeviceTPCH(a.sf)
    dcols = dg.columns("lineitem", needed(man, "li"))
    host = host_reference_columns(dcols, pin=True)
    del dcols, dg
    runtime.STORE.clear()
    torch.cuda.empty_cache()
    be = runtime.backend()
    rep_of = {c: r for arg, c, r in man["inputs"]}
    rows = []
    for name, r in rep_of.items():
        src = host[name]
        nbytes = src.nbytes
        # bare copy: pinned host -> a preallocated device buffer
        dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        hs = torch.from_numpy(src.view("uint8") if src.dtype.kind != "U" else src.view("uint32").view("uint8"))
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(hs, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        del dst
        bare = min(ts)
        runtime.STORE.enabled = False
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            col = runtime.STORE.get(src, r, 1)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            del col
        runtime.STORE.enabled = True
        row = {"column": name, "rep": r, "dtype": str(src.dtype), "bytes": nbytes, "pinned": bool(hs.is_pinned()),
               "bare_copy_ms": round(bare * 1e3, 2), "bare_GBps": round(nbytes / bare / 1e9, 1),
               "store_ms": round(min(ts) * 1e3, 2), "store_GBps": round(nbytes / min(ts) / 1e9, 1)}
        print(json.dumps(row), flush=True)
        rows.append(row)
    # the whole call
    hdb = reference_db(host)
    runtime.STORE.enabled = False
    fn = getattr(mod, a.query + "_compiled")
    fn(hdb)
    ts = []
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(hdb)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    tot = {"call_ms": round(min(ts) * 1e3, 2), "h2d_bytes": int(mod.last.h2d_bytes), "sum_store_ms": round(sum(r["store_ms"] for r in rows), 2),
           "sum_bare_ms": round(sum(r["bare_copy_ms"] for r in rows), 2), "device_ms": float(mod.last.device_ms)}
    print(json.dumps(tot), flush=True)
    if a.out:
        json.dump({"sf": a.sf, "query": a.query, "columns": rows, "total": tot}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
