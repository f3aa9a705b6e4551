#!/usr/bin/env python3
"""Run TPC-H queries on cuda:0 at a scale factor: device time per query, per-kernel times, optional comparison with
the reference module (oracle/_ref) on the same inputs.   python tools/run_tpch.py --sf 1 [--queries q1,q6] [--check]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=1.0)
    ap.add_argument("--queries", default=",".join("q%d" % i for i in range(1, 23)))
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ref-threads", type=int, default=os.cpu_count())
    ap.add_argument("--out", default=None)
    ap.add_argument("--device-gen", action="store_true", help="lineitem / orders generated on the GPU (SF100 on one GPU)")
    a = ap.parse_args()
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import runtime
    mod = runtime.load_compiled(os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py"))
    g = TPCH(a.sf)
    dg = None
    if a.device_gen:
        from sdqlpy_b200.tpch.gen_device import DeviceTPCH
        dg = DeviceTPCH(a.sf)
    ref = None
    if a.check:
        os.environ["SDQL_REF_THREADS"] = str(a.ref_threads)
        name = "tpchref_sf10_t8" if a.sf > 1 else "tpchref_sf1_t8"
        ref = rr.load(name)
    tabs, reftabs = {}, {}
    report = []
    for q in a.queries.split(","):
        man = mod.queries[q]
        db = []
        for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
            need = sorted({c for aa, c, r in man["inputs"] if aa == arg} |
                          {x.split(":")[3] for _, x in man["result"] if x.startswith("str:") and x.split(":")[2] == arg and len(x.split(":")) > 3})
            key = (t, tuple(need))
            if dg is not None and t in ("lineitem", "orders"):  # regenerated per query (fast), never cached: HBM stays free
                cols = dg.columns(t, need + [SCHEMAS[t][0][0]])
                db.append([cols.get(c) for c, _ in SCHEMAS[t]])
                continue
            if key not in tabs:
                cols = g.columns(t, need + [SCHEMAS[t][0][0]])
                tabs[key] = [cols.get(c) for c, _ in SCHEMAS[t]]
            db.append(tabs[key])
        t0 = time.time()
        res = mod.run(q, db)
        first = time.time() - t0
        ms, kms = [], []
        args_, keep = mod.prepare(q, db)
        for _ in range(a.reps):
            mod.execute(q, args_, fetch=False, kernel_times=True)
            ms.append(float(args_.device_ms))
            kms.append([args_.kernel_ms[k] for k in range(int(args_.launches))])
        kavg = np.mean(np.array(kms), axis=0)
        # algorithmic scan bytes: every relation-scan kernel reads its scanned columns once, in the
        # resident layout; table builds / probes / string-pattern columns are NOT counted (they lower the fraction)
        nrows = {arg: int(args_.nrows[i]) for i, arg in enumerate(man["args"])}
        wid = {"i32": 4, "f64": 8}
        scan_bytes = 0
        for k in man["kernels"]:
            if k["source"][0] == "rel":
                scan_bytes += nrows[k["source"][1]] * sum(wid.get(rep, 1) for _, rep in k["scan_cols"])
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        row = {"query": q, "sf": a.sf, "device_ms_min": min(ms), "device_ms_mean": float(np.mean(ms)),
               "scan_bytes": scan_bytes, "scan_GBps": round(scan_bytes / (min(ms) * 1e-3) / 1e9, 1),
               "frac_of_measured_hbm": round(scan_bytes / (min(ms) * 1e-3) / 1e9 / peak, 3),
               "first_call_s": first, "launches": int(args_.launches), "workspace_MB": mod.last.workspace_bytes / 1e6,
               "rows": mod.last.rows, "kernels": {man["kernels"][k]["name"]: round(float(kavg[k]), 4) for k in range(len(kavg))}}
        if ref is not None:
            rdb = []
            for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
                need = sorted({c for aa, c, r in man["inputs"] if aa == arg} |
                              {x.split(":")[3] for _, x in man["result"] if x.startswith("str:") and x.split(":")[2] == arg and len(x.split(":")) > 3})
                key = (t, tuple(need))
                if key not in reftabs:
                    reftabs[key] = g.ref_table(t, need)
                rdb.append(reftabs[key])
            rr.run(ref, q, rdb)
            t0 = time.time()
            want = rr.run(ref, q, rdb)
            row["ref_ms"] = (time.time() - t0) * 1e3
            row["ref_threads"] = a.ref_threads
            d = compare(res, want)
            if d is not None:  # the threaded reference's dense bool sets race (ref_runner.check): look again
                d, row["ref_runs"] = rr.check(ref, q, rdb, res, compare)
            row["parity"] = "ok" if d is None else d[:200]
        print(json.dumps(row), flush=True)
        report.append(row)
        del db, args_, keep, res
        if dg is not None:
            runtime.STORE.clear()
            mod.ws, mod.ws_bytes = None, 0   # per-query workspace: the next query may need a very different size
            import torch
            torch.cuda.empty_cache()
    if a.out:
        json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
