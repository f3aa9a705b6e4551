#!/usr/bin/env python3
"""TPC-H queries on N GPUs of one node (one process per GPU, NCCL): lineitem/orders range partitioned on order
boundaries, dimension tables replicated, partial tables merged through the C-ABI merge callback.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_tpch_dist.py --sf 100 --queries q9,q18
  (N = 1 works without torchrun)

--check ref : rank 0 also runs the reference module (oracle/_ref, all host threads) on the full data set and compares
--check port: rank 0 runs the pandas restatement (oracle/tpch_port.py) instead
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH  # noqa: E402


def needed(man, arg):
    cols = {c for a, c, r in man["inputs"] if a == arg}
    for _, k in man["result"]:
        f = k.split(":")
        if f[0] == "str" and f[1] in ("ref", "code") and f[2] == arg:
            cols.add(f[3])
    return sorted(cols)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10)
    ap.add_argument("--queries", default="q1,q6,q3,q5,q9,q18")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", default="none")
    ap.add_argument("--out", default=None)
    ap.add_argument("--device-gen", action="store_true", help="lineitem / orders partitions generated on the GPU")
    ap.add_argument("--trace", action="store_true", help="one extra run per query with SDQLB200_F_TRACE: rank 0 reports the "
                    "wall-clock time of every step of the host driver (kernels, table initialisation, merges)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import runtime
    if world > 1:
        runtime.set_distributed(runtime.DistConfig(partitioned=("li", "ord")))
    mod = runtime.load_compiled(os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py"))
    g = TPCH(a.sf)
    dg = None
    if a.device_gen:
        from sdqlpy_b200.tpch.gen_device import DeviceTPCH
        dg = DeviceTPCH(a.sf)
    per = g.O // world
    orng = (rank * per, (rank + 1) * per if rank < world - 1 else g.O)
    cache = {}
    report = []
    for q in a.queries.split(","):
        man = mod.queries[q]
        db = []
        t0 = time.time()
        for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
            fact = t in ("lineitem", "orders")
            missing = [c for c in needed(man, arg) if (t, c) not in cache]
            if missing:
                src = dg if (fact and dg is not None) else g
                for c, col in src.columns(t, missing, orng if fact else None).items():
                    cache[(t, c)] = col
            db.append([cache.get((t, c)) for c, _ in SCHEMAS[t]])
        gen_s = time.time() - t0
        try:
            res = mod.run(q, db)          # first call: upload + run
        except Exception as e:  # noqa: BLE001 -- report and go on with the next query (all ranks fail alike)
            if rank == 0:
                print(json.dumps({"query": q, "sf": a.sf, "n_gpus": world, "error": repr(e)[:300]}), flush=True)
            runtime.STORE.clear()
            cache.clear()
            continue
        dev, wall = [], []
        for _ in range(a.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = mod.run(q, db)
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            dev.append(mod.last.device_ms)
        t = torch.tensor([min(wall), min(dev)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows = res.tuples() if hasattr(res, "tuples") else res
        row = {"query": q, "sf": a.sf, "n_gpus": world, "latency_ms_wall": float(t[0]), "device_ms": float(t[1]),
               "gen_s": round(gen_s, 1), "result_rows": len(rows) if isinstance(rows, list) else 1,
               "workspace_MB": round(mod.last.workspace_bytes / 1e6, 1), "merges_total": mod.merges,
               "table_merges_total": mod.table_merges, "p2p_merges_total": mod.p2p_merges}
        if a.trace:  # every rank runs it (the merges are collective); rank 0 keeps the lines
            from run_tpch import traced_run
            args_, keep_ = mod.prepare(q, db)
            steps = traced_run(mod, q, args_)
            row["trace_ms"] = [[s_, round(ms_, 4)] for s_, ms_ in steps if ms_ >= 0.02]
            del args_, keep_
        if rank == 0 and a.check != "none":
            full = TPCH(a.sf)
            t0 = time.time()
            if a.check == "ref":
                os.environ["SDQL_REF_THREADS"] = str(os.cpu_count())
                ref = rr.load("tpchref_sf10_t8" if a.sf <= 10 else "tpchref_sf100_t8")
                if world == 1:  # the rank already holds the full data set: convert instead of regenerating
                    rdb = []
                    for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
                        want_cols = set(needed(man, arg)) | {SCHEMAS[t][0][0]}
                        tab = []
                        for c, kind in SCHEMAS[t]:
                            if c in want_cols:
                                if (t, c) not in cache:
                                    cache[(t, c)] = g.columns(t, [c])[c]
                                tab.append(np.ascontiguousarray(cache[(t, c)].to_ref()))
                            elif isinstance(kind, tuple):
                                tab.append(np.zeros(1, dtype="<U%d" % kind[1]))
                            else:
                                tab.append(np.zeros(1, dtype=np.float64 if kind == "float" else np.int64))
                        rdb.append(tab)
                else:
                    rdb = [full.ref_table(t, needed(man, arg)) for arg, t in zip(man["args"], rr.QUERY_ARGS[q])]
                row["check_gen_s"] = round(time.time() - t0, 1)
                t0 = time.time()
                want = None
                d, runs = rr.check(ref, q, rdb, res, compare)
                row["ref_runs"] = runs
            else:
                import tpch_port
                pdb = {}
                for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
                    cols = full.columns(t, needed(man, arg))
                    pdb[t] = {c: v.to_ref() for c, v in cols.items()}
                row["check_gen_s"] = round(time.time() - t0, 1)
                t0 = time.time()
                want = tpch_port.QUERIES[q](pdb)
            row["check_ms"] = round((time.time() - t0) * 1e3, 1)
            if want is not None:
                d = compare(res, want)
            row["parity_vs_" + a.check] = "ok" if d is None else d[:300]
        if rank == 0:
            if isinstance(rows, list):
                row["sample"] = repr(sorted(rows, key=repr)[:2])[:300]
            else:
                row["sample"] = repr(rows)
            print(json.dumps(row), flush=True)
            report.append(row)
        if world > 1:
            dist.barrier()
        if a.sf >= 30:  # keep host + device memory bounded at large scale: drop this query's fact-table columns
            if dg is None:
                runtime.STORE.clear()
                cache.clear()
            else:  # device-generated partitions are regenerated per query; host dimension tables stay cached
                for k in [k for k in cache if k[0] in ("lineitem", "orders")]:
                    del cache[k]
            del db, res
            mod.ws, mod.ws_bytes = None, 0
            torch.cuda.empty_cache()
    if rank == 0 and a.out:
        json.dump(report, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
