#!/usr/bin/env python3
"""Headline benchmark: TPC-H Q1 at SF10 per GPU (BASELINE.json configs[1]) -- scan GB/s and per-query latency.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--query q1] [--sf 10]

One "step" = one execution of the query over one batch of synthetic lineitem rows.
  value   : whole-job scan throughput, algorithmic bytes of the DEVICE layout / CUDA-event time, inputs resident in HBM
  e2e     : the same metric through the reference-facing call <fn>_compiled(db) with HOST (pinned) columns: every step
            re-uploads the query's input columns, runs, and reads the result back
  roofline: dominant kernel, bytes per launch / its own CUDA-event time, against MEASURED_PEAKS.json
  N > 1   : weak scaling -- rank r owns the r-th order range of an SF*N database (lineitem range partitioned on order
            boundaries, as north_star), partial results are exchanged with one NCCL all-gather per step and merged.
--impl reference times the reference's own generated C++ (oracle/_ref, TBB-shim threads = host cores) on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from sdqlpy_b200 import wire  # noqa: E402
from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH  # noqa: E402

QUERY_SCRIPT = os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py")
ELEM_BYTES = {"i32": 4, "f64": 8}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML in a background thread DURING the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu=0):
        self.gpu, self.sm, self.mask, self.stop_flag, self.th, self.h = gpu, [], 0, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def start(self):
        if self.h is not None:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.th.join()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": float(self.max),
                "samples": len(self.sm), "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k)}


def scan_bytes_per_row(man, relation_arg):
    """algorithmic bytes one scanned row of ``relation_arg`` costs in the device layout."""
    total, cols = 0, []
    for k in man["kernels"]:
        if k["source"] == ["rel", relation_arg]:
            for c, rep in k["scan_cols"]:
                b = ELEM_BYTES.get(rep, 1)
                total += b
                cols.append("%s:%d" % (c, b))
    return total, cols


def lineitem_columns(g, man, order_range):
    need = sorted({c for a, c, r in man["inputs"] if a == "li"})
    return g.columns("lineitem", need, order_range)


def result_check(q, cols, result):
    """size-independent properties of the bench query's result at the full bench size, against numpy reductions over the
    host columns (outside every timed region): row counts and integer-valued sums exactly, fp64 sums within 1e-9
    relative.  -> "ok" or a description of the first violated property (reported in the JSON line, never raised)."""
    try:
        ship = np.asarray(cols["l_shipdate"].data)
        if q == "q1":
            m = ship <= 19980902
            rows = result.tuples()
            if sum(r[-1] for r in rows) != int(m.sum()):
                return "sum of count_order %d != %d qualifying rows" % (sum(r[-1] for r in rows), int(m.sum()))
            qty = float(np.asarray(cols["l_quantity"].data)[m].sum())  # integer-valued: exact in any order
            if sum(r[2] for r in rows) != qty:
                return "sum of sum_qty %r != %r" % (sum(r[2] for r in rows), qty)
            base = float(np.asarray(cols["l_extendedprice"].data)[m].sum())
            got = sum(r[3] for r in rows)
            if abs(got - base) > 1e-9 * abs(base):
                return "sum of sum_base_price %r != %r" % (got, base)
            groups = len(set(zip(np.asarray(cols["l_returnflag"].data)[m].tolist(), np.asarray(cols["l_linestatus"].data)[m].tolist())))
            if len(rows) != groups:
                return "%d groups != %d" % (len(rows), groups)
            return "ok"
        if q == "q6":
            disc, qty = np.asarray(cols["l_discount"].data), np.asarray(cols["l_quantity"].data)
            m = (ship >= 19940101) & (ship < 19950101) & (disc >= 0.05) & (disc <= 0.07) & (qty < 24.0)
            want = float((np.asarray(cols["l_extendedprice"].data)[m] * disc[m]).sum())
            return "ok" if abs(float(result) - want) <= 1e-9 * abs(want) else "revenue %r != %r" % (float(result), want)
        return "no check for %s" % q
    except Exception as ex:  # a broken checker must not cost the bench line
        return "checker failed: %r" % (ex,)


def ref_arm(args, nproc):
    """the reference's own CPU implementation of the path, all host threads, on a bounded sample."""
    import ref_runner as rr
    os.environ["SDQL_REF_THREADS"] = str(nproc)
    name = "tpchref_sf10_t8" if rr.available("tpchref_sf10_t8") else "tpchref_sf1_t8"
    mod = rr.load(name)
    sample_sf = min(args.sf, args.ref_sample_sf)
    g = TPCH(sample_sf)
    q = args.query
    db = []
    man = json.load(open(os.path.join(ROOT, "sdqlpy_b200", "tpch", "sdqlb200_generated", "manifest.json")))
    qm = [x for x in man["queries"] if x["name"] == q][0]
    for a, t in zip(qm["args"], rr.QUERY_ARGS[q]):
        db.append(g.ref_table(t, [c for aa, c, r in qm["inputs"] if aa == a]))
    rows = len(db[qm["args"].index("li")][0]) if "li" in qm["args"] else 0
    bpr, _ = scan_bytes_per_row(qm, "li")
    fn = getattr(mod, q + "_compiled")
    for _ in range(args.warmup):
        fn(db)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(db)
    dt = (time.perf_counter() - t0) / args.steps
    gbs = rows * bpr / dt / 1e9
    sample = "lineitem of TPC-H SF%g (%d rows, reference layout int64/fp64/UCS4), %s, %d TBB-shim threads" % (
        sample_sf, rows, name, nproc)
    return gbs, dt * 1e3, rows, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--query", default="q1")
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--ref-sample-sf", type=float, default=2.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nproc = os.cpu_count() or 1
    q = args.query
    workload = "tpch_%s_sf%g_per_gpu" % (q, args.sf)

    if args.impl == "reference":
        if rank != 0:
            return
        gbs, ms, rows, sample = ref_arm(args, nproc)
        print(json.dumps({
            "impl": "reference", "metric": "tpch_%s_scan_throughput" % q, "value": gbs, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "rows_per_step": rows,
                       "bytes_per_row": "device-layout bytes (same numerator as the b200 arm)"},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": nproc, "kind": "reference", "sample": sample},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sdqlpy_b200 import runtime
    if world > 1:  # lineitem is range partitioned on order boundaries across the ranks
        runtime.set_distributed(runtime.DistConfig(partitioned=("li",), partkeys=("l_orderkey",)))
    mod = runtime.load_compiled(QUERY_SCRIPT)
    man = mod.queries[q]
    if man["args"] != ["li"]:
        raise SystemExit("bench.py drives single-relation lineitem scans (q1, q6); use tools/run_tpch.py for the others")
    # ---- data: rank r owns the r-th order range of an SF*world database -------------------------------------
    g = TPCH(args.sf * world)
    o_per = g.O // world
    orng = (rank * o_per, (rank + 1) * o_per if rank < world - 1 else g.O)
    cols = lineitem_columns(g, man, orng)
    names = [c for c, _ in SCHEMAS["lineitem"]]
    be = runtime.backend()
    keep, wire_cols, wire_bpr = [], [], 0
    t_pack = time.perf_counter()
    for c, col in cols.items():
        # load time (what read_csv is to the reference): every column gets its lossless packed image for the
        # host -> device link (sdqlpy_b200/wire.py); host copies live in page-locked memory (source of the e2e uploads)
        wire.pack_column(col)
        if col.wire is not None:
            col.wire.pin(be)
            bits = col.wire.nbits if col.wire.nbits else 8 * col.wire.codes.itemsize
            wire_cols.append("%s:%s:%db" % (c, wire.KIND_NAMES[col.wire.kind], bits))
            wire_bpr += bits / 8.0
        else:
            v, t = be.pinned_like(col.data)
            col.data = v
            keep.append(t)
            wire_cols.append("%s:plain:%db" % (c, 8 * col.data.itemsize))
            wire_bpr += col.data.itemsize
    t_pack = time.perf_counter() - t_pack
    db = [[cols.get(c) for c in names]]
    rows = len(next(iter(cols.values())).data)
    bpr, bcols = scan_bytes_per_row(man, "li")
    step_bytes = rows * bpr
    # ---- device-resident timing ---------------------------------------------------------------------------
    a, keepalive = mod.prepare(q, db)
    torch.cuda.synchronize()

    # N > 1: partial tables were merged inside the query (all-reduce through the merge callback); every group is emitted
    # by its owner rank, so the ranks' result rows are concatenated with one NCCL all-gather per step.  The rows go
    # through a preallocated pinned buffer: no allocation and no extra synchronisation per step (the next step's result
    # fetch synchronises the stream before the buffer is written again).
    ct = __import__("ctypes")
    MAXR, MAXF = 64, 8
    if world > 1:
        pin = torch.zeros(MAXR, MAXF, dtype=torch.int64, pin_memory=True)
        pin_np = pin.numpy()
        dbuf = torch.empty(MAXR, MAXF, dtype=torch.int64, device="cuda")
        gathered = torch.empty(world * MAXR, MAXF, dtype=torch.int64, device="cuda")

    def step():
        mod.execute(q, a, fetch=True)
        res = a.result
        n, nf = min(int(res.count), MAXR - 1), min(int(res.nfields), MAXF)
        if world > 1:
            pin_np[MAXR - 1, 0] = n  # rows this rank contributes
            for j in range(nf):
                pin_np[:n, j] = np.ctypeslib.as_array(res.cols[j], shape=(max(n, 1),))[:n]
            dbuf.copy_(pin, non_blocking=True)
            dist.all_gather_into_tensor(gathered, dbuf)
        mod.lib.sdqlb200_result_free(ct.byref(a.result))
        return float(a.device_ms), int(a.launches)

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms, launches = [], 0
    e0.record()
    for _ in range(args.steps):
        ms, ln = step()
        dev_ms.append(ms)
        launches += ln
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    rows_t = torch.tensor([rows], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(rows_t)
    all_bytes = float(rows_t.item()) * bpr
    value = all_bytes / (ms_per_step * 1e-3) / 1e9
    # ---- roofline of the dominant kernel (its own CUDA events) ---------------------------------------------
    kms = []
    for _ in range(5):
        mod.execute(q, a, fetch=True, kernel_times=True)
        mod.lib.sdqlb200_result_free(__import__("ctypes").byref(a.result))
        kms.append([a.kernel_ms[k] for k in range(int(a.launches))])
    kavg = np.mean(np.array(kms), axis=0)
    dom = int(np.argmax(kavg))
    peak, peak_src = peaks()
    achieved = step_bytes / (kavg[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": man["kernels"][dom]["name"], "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": float(kavg[dom]), "bytes_per_launch": step_bytes,
                "kernel_share_of_step": float(kavg[dom] / max(1e-9, np.mean(dev_ms)))}
    prof = os.path.join(ROOT, "profiles", "r01_%s_traffic.json" % q)
    if os.path.exists(prof):
        roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
    # ---- end to end through <fn>_compiled(db): host columns, upload every step -------------------------------
    runtime.STORE.enabled = False
    runtime.STORE.clear()
    fn = getattr(mod, q + "_compiled")
    r = fn(db)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        r = fn(db)  # N > 1: CompiledModule.run gathers the ranks' result rows itself
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": all_bytes / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": int(mod.last.h2d_bytes), "d2h_bytes_per_step": int(mod.last.d2h_bytes),
           "wire_bytes_per_row": wire_bpr, "wire_layout": wire_cols, "pack_s_at_load": round(t_pack, 2),
           "note": "every step: packed host columns (pinned) -> PCIe -> device expansion to the resident layout -> "
                   "query -> result to host; %d steps; value = resident-layout bytes / time (same numerator as "
                   "'value' and as the reference arm)" % args.e2e_steps}
    runtime.STORE.enabled = True
    checked = result_check(q, cols, r) if world == 1 else "not checked (N > 1: the host columns are this rank's partition)"
    out = {
        "metric": "tpch_%s_scan_throughput" % q, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "query": q, "sf_per_gpu": args.sf, "rows_per_gpu": rows,
                   "bytes_per_row": bpr, "layout": bcols, "l2_policy": "inputs (%.2f GB per GPU) larger than L2" % (step_bytes / 1e9),
                   "latency_ms_device": float(np.mean(dev_ms)), "agg_tier": int(a.tier)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "result_check": checked,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            a2 = argparse.Namespace(**vars(args))
            a2.steps, a2.warmup = 5, 1
            gbs, ms, rrows, sample = ref_arm(a2, nproc)
            out["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": nproc, "kind": "reference", "sample": sample,
                                   "ms_per_step": ms}
        except Exception as ex:  # the checker is missing: say so, do not fake a number
            out["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": nproc, "kind": "reference",
                                   "sample": "unavailable: %r" % (ex,)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
