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


def traced_run(mod, q, args_):
    """one execution with SDQLB200_F_TRACE; the driver's per-step lines on stderr are captured -> [[step, ms], ...]"""
    import tempfile
    sys.stderr.flush()
    with tempfile.TemporaryFile(mode="w+b") as tf:
        saved = os.dup(2)
        os.dup2(tf.fileno(), 2)
        try:
            mod.execute(q, args_, fetch=False, trace=True)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
        tf.seek(0)
        text = tf.read().decode("latin1")
    out = []
    for line in text.splitlines():
        p = line.split()
        if len(p) >= 5 and p[0] == "[sdqlb200]" and p[1] == "step":
            out.append([p[2], float(p[3])])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=1.0)
    ap.add_argument("--queries", default=",".join("q%d" % i for i in range(1, 23)))
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ref-threads", type=int, default=os.cpu_count())
    ap.add_argument("--out", default=None)
    ap.add_argument("--device-gen", action="store_true", help="lineitem / orders generated on the GPU (SF100 on one GPU)")
    ap.add_argument("--trace", action="store_true", help="one extra run per query with SDQLB200_F_TRACE: wall-clock time "
                    "of every step of the host driver (count pass, table initialisation, kernel, merge, presence bits)")
    ap.add_argument("--stats-so", default=None, help="counting build of the module (tools/build_variant.py stats ONLY=all "
                    "NVCCDEF=SDQLB200_STATS): run once per query, untimed, for the bytes-moved roofline")
    a = ap.parse_args()
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import roofline, runtime
    mod = runtime.load_compiled(os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py"))
    smod = runtime.CompiledModule(a.stats_so) if a.stats_so else None
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
        steps = None
        if a.trace:
            steps = traced_run(mod, q, args_)
        # algorithmic scan bytes (sdqlpy_b200/roofline.py): every relation-scan kernel reads its streamed
        # columns (and the bytes of the string columns it searches) once, in the resident layout; table builds / probes
        # are NOT counted here (they lower this fraction) -- they are what the bytes-moved figure below adds
        nrows = {arg: int(args_.nrows[i]) for i, arg in enumerate(man["args"])}
        col_b, str_b = roofline.scan_bytes(man, nrows)
        scan_bytes = col_b + str_b
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        row = {"query": q, "sf": a.sf, "device_ms_min": min(ms), "device_ms_mean": float(np.mean(ms)),
               "scan_bytes": scan_bytes, "scan_string_bytes": str_b, "scan_GBps": round(scan_bytes / (min(ms) * 1e-3) / 1e9, 1),
               "frac_of_measured_hbm": round(scan_bytes / (min(ms) * 1e-3) / 1e9 / peak, 3),
               "first_call_s": first, "launches": int(args_.launches), "workspace_MB": mod.last.workspace_bytes / 1e6,
               "rows": mod.last.rows, "kernels": {man["kernels"][k]["name"]: round(float(kavg[k]), 4) for k in range(len(kavg))}}
        if steps is not None:
            row["steps_ms"] = steps
        if smod is not None and q in smod.queries:
            # counting build, same device-resident inputs, one untimed run: data-dependent accesses of the query
            smod.ws, smod.ws_bytes = None, 0
            del args_
            args_ = keep = None
            mod.ws, mod.ws_bytes = None, 0        # the two modules never hold a workspace at the same time (SF100)
            if runtime.backend().name != "emu":
                import torch
                torch.cuda.empty_cache()
            sargs, skeep = smod.prepare(q, db)
            smod.execute(q, sargs, fetch=False)   # sizes the workspace (dry run + first run)
            smod.stats()
            smod.execute(q, sargs, fetch=False)
            st, counting = smod.stats()
            if counting:
                bm = roofline.bytes_moved(man, nrows, st, mod.last.rows or 0)
                row["stats"] = st
                row["bytes_moved"] = bm
                row["moved_GBps"] = round(bm["total"] / (min(ms) * 1e-3) / 1e9, 1)
                row["frac_bytes_moved_of_measured_hbm"] = round(bm["total"] / (min(ms) * 1e-3) / 1e9 / peak, 3)
            del sargs, skeep
            smod.ws, smod.ws_bytes = None, 0
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
