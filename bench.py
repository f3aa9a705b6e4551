#!/usr/bin/env python3
"""Headline benchmark: TPC-H at SF100 (BASELINE.json metric), fact tables generated in HBM, STRONG-scaled over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--sf 100] [--queries all|q1,q9,..]

One "step" = one execution of Q1 (filtered scan + low-cardinality group-by) over the lineitem rows of the whole SF100
database: 600 M rows, range partitioned over the N ranks on order boundaries (north_star), partial group tables merged over
NVLink (csrc/sdqlb200_comm.cu) and the result rows brought to the host, every step.
  value        : whole-job Q1 scan throughput = algorithmic bytes of the DEVICE layout (38 B/row) / CUDA-event time, max over
                 ranks, inputs resident in HBM
  roofline     : q1_k0 (dominant kernel) and q6_k0: bytes per launch / the kernel's own CUDA-event time vs MEASURED_PEAKS.json
  per_query_ms : all 22 queries at SF100 (device time: every kernel, table initialisation and merge of the query; min of 3
                 runs, max over ranks), each with a parity verdict (fingerprints of the REFERENCE's results where committed,
                 tests/golden/tpch_sf*_fingerprints.json; exact torch reductions for Q1 / Q6)
  e2e          : Q1 through the reference-facing call <fn>_compiled(db) with plain reference-layout numpy columns in HOST
                 memory (int64 / float64 / <U1, what the reference's db contract passes, sdql_lib.py:420-424): every step
                 uploads the raw columns, converts them on the device, runs the query and reads the result back
  cpu_baseline : the reference's own generated C++ (oracle/_ref, TBB-shim threads = host cores) on the SAME host columns
--impl reference times that reference module alone (rank 0; whole SF100 lineitem per step when the box has the memory).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH  # noqa: E402

QUERY_SCRIPT = os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py")
ELEM_BYTES = {"i32": 4, "f64": 8}
ALL22 = ["q%d" % i for i in range(1, 23)]
DIMS = ("supplier", "customer", "part", "partsupp", "nation", "region")


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


def needed(man, arg):
    cols = {c for a, c, r in man["inputs"] if a == arg}
    for _, k in man["result"]:
        f = k.split(":")
        if f[0] == "str" and f[1] in ("ref", "code") and len(f) > 3 and f[2] == arg:
            cols.add(f[3])
    return sorted(cols)


def result_check(q, cols, result):
    """size-independent properties of the bench query's result, against numpy reductions over host columns (outside every
    timed region): row counts and integer-valued sums exactly, fp64 sums within 1e-9 relative.  -> "ok" or a description
    of the first violated property (reported in the JSON line, never raised)."""
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


def device_check(q, dcols, result, allsum):
    """the same properties at full size, from torch reductions over this rank's DEVICE columns (the checker: independent of
    the generated kernels), summed over the ranks by ``allsum``"""
    try:
        import torch

        def t(name):
            return dcols[name].holder[:dcols[name].rows]
        ship = t("l_shipdate")
        if q == "q1":
            m = ship <= 19980902
            cnt, qty = allsum([float(m.sum()), float(t("l_quantity")[m].sum())])
            base = allsum([float(t("l_extendedprice")[m].sum(dtype=torch.float64))])[0]
            rows = result.tuples()
            if sum(r[-1] for r in rows) != int(cnt):
                return "sum of count_order %d != %d qualifying rows" % (sum(r[-1] for r in rows), int(cnt))
            if sum(r[2] for r in rows) != qty:
                return "sum of sum_qty %r != %r" % (sum(r[2] for r in rows), qty)
            got = sum(r[3] for r in rows)
            if abs(got - base) > 1e-9 * abs(base):
                return "sum of sum_base_price %r != %r" % (got, base)
            return "ok"
        if q == "q6":
            disc, qty = t("l_discount"), t("l_quantity")
            m = (ship >= 19940101) & (ship < 19950101) & (disc >= 0.05) & (disc <= 0.07) & (qty < 24.0)
            want = allsum([float((t("l_extendedprice")[m] * disc[m]).sum(dtype=torch.float64))])[0]
            return "ok" if abs(float(result) - want) <= 1e-9 * abs(want) else "revenue %r != %r" % (float(result), want)
        return None
    except Exception as ex:  # noqa: BLE001
        return "checker failed: %r" % (ex,)


def host_reference_columns(dcols, pin):
    """this rank's device-resident Q1 columns -> numpy columns in the REFERENCE layout (int64 / float64 / <U1) in host
    memory (page-locked when ``pin``): what read_csv would have produced (sdql_lib.py:83-97)."""
    import torch
    out = {}
    for name, dc in dcols.items():
        n = dc.rows
        if dc.kind == "i32":
            h = torch.empty(n, dtype=torch.int64, pin_memory=pin)
            h.copy_(dc.holder[:n])      # int32 -> int64 on the way
            out[name] = h.numpy()
        elif dc.kind == "f64":
            h = torch.empty(n, dtype=torch.float64, pin_memory=pin)
            h.copy_(dc.holder[:n])
            out[name] = h.numpy()
        else:                           # dictionary code -> the string itself, UCS4 like numpy's <U1
            assert all(len(s) == 1 for s in dc.dictionary), "only <U1 code columns are expanded here"
            lut = torch.tensor([ord(s) for s in dc.dictionary], dtype=torch.int32, device=dc.holder.device)
            h = torch.empty(n, dtype=torch.int32, pin_memory=pin)
            h.copy_(lut[dc.holder[:n].long()])
            out[name] = h.numpy().view("<U1")
        out[name + "__keep"] = h
    torch.cuda.synchronize()
    return out


def reference_db(host_cols):
    """the reference's db for a lineitem-only query: all 17 columns in schema order, 1-element placeholders for the ones the
    query does not read (the reference casts every pointer and reads only what the query uses, sdql_compiler.py:652-668)"""
    rel = []
    for c, k in SCHEMAS["lineitem"]:
        if c in host_cols:
            rel.append(host_cols[c])
        elif c == "l_orderkey":  # column 0 carries the row count (sdql_compiler.py:644): a zero-copy view of the right length
            n = len(host_cols["l_shipdate"])
            rel.append(np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int64), shape=(n,), strides=(0,)))
        elif isinstance(k, tuple):
            rel.append(np.zeros(1, dtype="<U%d" % k[1]))
        elif k == "float":
            rel.append(np.zeros(1, dtype=np.float64))
        else:
            rel.append(np.zeros(1, dtype=np.int64))
    return [rel]


def time_reference(q, db, rows, bpr, steps, warmup, nproc, sf):
    import ref_runner as rr
    os.environ["SDQL_REF_THREADS"] = str(nproc)
    name = next((n for n in ("tpchref_sf100_t8", "tpchref_sf10_t8", "tpchref_sf1_t8") if rr.available(n)), None)
    if name is None:
        raise RuntimeError("oracle/_ref is not built")
    fn = getattr(rr.load(name), q + "_compiled")
    for _ in range(warmup):
        fn(db)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = fn(db)
    dt = (time.perf_counter() - t0) / steps
    sample = "lineitem of TPC-H SF%g (%d rows, reference layout int64/fp64/UCS4 = %d B/row read), module %s, %d TBB-shim threads" % (
        sf, rows, 48 if q == "q1" else 32, name, nproc)
    return rows * bpr / dt / 1e9, dt * 1e3, sample, rr.normalise(res)


def lineitem_host_columns(sf, man, steps_hint):
    """host columns of the whole SF lineitem for the reference arm: generated on the GPU when there is one (the generator
    only -- nothing of the timed path), else by the numpy generator on a bounded sample"""
    need = needed(man, "li")
    try:
        import torch
        if torch.cuda.is_available():
            from sdqlpy_b200 import runtime
            from sdqlpy_b200.tpch.gen_device import DeviceTPCH
            free = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
            use = sf
            while use > 1 and 60e6 * use / 10 * 60 > 0.6 * free:  # ~48 B/row + staging
                use /= 2
            dg = DeviceTPCH(use)
            dcols = dg.columns("lineitem", need)
            host = host_reference_columns(dcols, pin=False)
            del dcols, dg
            runtime.STORE.clear()
            torch.cuda.empty_cache()
            return host, use, "generated on the GPU, copied to the host"
    except Exception as ex:  # noqa: BLE001
        sys.stderr.write("bench: device generation unavailable (%r), numpy generator\n" % (ex,))
    use = min(sf, 10.0)
    g = TPCH(use)
    cols = g.ref_table("lineitem", need)
    return {c: cols[i] for i, (c, _) in enumerate(SCHEMAS["lineitem"]) if c in need}, use, "numpy generator (no GPU)"


def reference_arm(args, nproc, workload):
    man = json.load(open(os.path.join(ROOT, "sdqlpy_b200", "tpch", "sdqlb200_generated", "manifest.json")))
    qm = [x for x in man["queries"] if x["name"] == args.query][0]
    bpr, bcols = scan_bytes_per_row(qm, "li")
    host, used_sf, how = lineitem_host_columns(args.sf, qm, args.steps)
    rows = len(host["l_shipdate"])
    gbs, ms, sample, _ = time_reference(args.query, reference_db(host), rows, bpr, args.steps, max(1, args.warmup), nproc, used_sf)
    return {
        "impl": "reference", "metric": "tpch_sf%g_%s_scan_throughput" % (args.sf, args.query), "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sf": args.sf, "query": args.query, "bytes_per_row": bpr},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": nproc, "kind": "reference",
                         "sample": sample + ("" if used_sf == args.sf else " [sample: SF%g of SF%g fits the host memory]" % (used_sf, args.sf)) + "; " + how},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rows_per_step": rows,
    }


def load_fingerprints(sf):
    tag = ("%g" % sf).replace(".", "p")
    p = os.path.join(ROOT, "tests", "golden", "tpch_sf%s_fingerprints.json" % tag)
    return json.load(open(p))["queries"] if os.path.exists(p) else {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--query", default="q1")
    ap.add_argument("--sf", type=float, default=100.0)
    ap.add_argument("--queries", default="all", help="queries of the per-query table: all | none | q1,q9,..")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nproc = os.cpu_count() or 1
    q = args.query
    workload = "tpch_sf%g" % args.sf

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, nproc, workload)))
        return

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's own lines (version banner, INFO) stay off stdout: ONE JSON line
    if world > 1 and os.environ.get("SDQLB200_BENCH_AFFINITY", "1") != "0":
        # one slice of the host cores per rank: the ranks meet in an all-reduce every step, so one rank's scheduling jitter is
        # everybody's step time (N = 8: 0.15 ms of a 0.69 ms step was spent outside the device events)
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
        except (AttributeError, OSError):
            pass
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ref_runner as rr
    from fingerprint import fingerprint, match
    from sdqlpy_b200 import runtime
    from sdqlpy_b200.tpch.gen_device import DeviceTPCH
    if world > 1:  # lineitem / orders range partitioned on order boundaries; NCCL id + peer mailboxes set up here
        runtime.set_distributed(runtime.DistConfig(partitioned=("li", "ord")))
    D = runtime.dist_config()
    mod = runtime.load_compiled(QUERY_SCRIPT)
    man = mod.queries[q]
    if man["args"] != ["li"]:
        raise SystemExit("the headline step is a single-relation lineitem scan (q1, q6)")
    qlist = [] if args.queries == "none" else (ALL22 if args.queries == "all" else args.queries.split(","))

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return [float(x) for x in t]

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- dimension tables: host generator, in the background (they are only needed by the per-query table) -------------
    g = TPCH(args.sf)
    dim_need = {}
    for qq in qlist:
        for arg, tname in zip(mod.queries[qq]["args"], rr.QUERY_ARGS[qq]):
            if tname in DIMS:
                dim_need.setdefault(tname, set()).update(needed(mod.queries[qq], arg) + [SCHEMAS[tname][0][0]])
    dims, dim_err = {}, []

    def gen_dim(tname):
        try:
            cols = g.columns(tname, sorted(dim_need[tname]))
            dims[tname] = [cols.get(c) for c, _ in SCHEMAS[tname]]
        except Exception as ex:  # noqa: BLE001
            dim_err.append((tname, ex))
    t_dims = time.perf_counter()
    dim_threads = [threading.Thread(target=gen_dim, args=(tname,), daemon=True) for tname in dim_need]
    for th in dim_threads:
        th.start()

    # ---- fact tables: rank r owns the r-th order range of the SF database, generated in HBM ------------------------------
    dg = DeviceTPCH(args.sf)
    o_per = g.O // world
    orng = (rank * o_per, (rank + 1) * o_per if rank < world - 1 else g.O)
    names = [c for c, _ in SCHEMAS["lineitem"]]
    dcols = dg.columns("lineitem", needed(man, "li"), orng)
    db = [[dcols.get(c) for c in names]]
    rows = next(iter(dcols.values())).rows
    rows_total = int(allsum([rows])[0])
    bpr, bcols = scan_bytes_per_row(man, "li")
    all_bytes = rows_total * bpr
    a, keepalive = mod.prepare(q, db)
    ct = __import__("ctypes")

    def step():
        """one Q1 over the whole database: kernels + merge on every rank, result rows on the host of every rank"""
        # returns with the result rows in host memory (a.result.cols, malloc'ed by the module); the module records a CUDA event
        # around every kernel on its own stream (kernel_times): the roofline's launch durations come from the timed steps
        mod.execute(q, a, fetch=True, kernel_times=True)
        if world > 1 and int(a.result_partial):  # rows emitted per owner rank: concatenated on every rank (not the case for Q1 / Q6)
            res = a.result
            n, nf = int(res.count), int(res.nfields)
            D.gather_rows([np.ctypeslib.as_array(res.cols[j], shape=(max(n, 1),))[:n] for j in range(nf)])
        mod.lib.sdqlb200_result_free(ct.byref(a.result))
        if ksum is not None:
            for k in range(int(a.launches)):
                ksum[k] += a.kernel_ms[k]
        return int(a.launches)

    W = max(3, args.warmup)
    ksum = None
    for _ in range(W):
        step()
    ksum = [0.0] * 24  # per-kernel CUDA-event time summed over the timed steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        ln = step()
        launches += ln + (3 if world > 1 else 0)  # + pack / peer all-reduce / unpack of the 6-slot group table
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_per_step = allmax(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = all_bytes / (ms_per_step * 1e-3) / 1e9
    # ---- rooflines: the kernels' own CUDA events (recorded around each launch inside the module) -------------------------
    peak, peak_src = peaks()

    def kernel_roofline(qq, aa, mm, kavg=None):
        if kavg is None:  # not the timed query: its own steps
            kms = []
            for _ in range(args.steps):
                mod.execute(qq, aa, fetch=True, kernel_times=True)
                mod.lib.sdqlb200_result_free(ct.byref(aa.result))
                kms.append([aa.kernel_ms[k] for k in range(int(aa.launches))])
            kavg = np.mean(np.array(kms), axis=0)
        dom = int(np.argmax(kavg))
        b1, _ = scan_bytes_per_row(mm, "li")
        kbytes = rows * b1
        ach = kbytes / (kavg[dom] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": mm["kernels"][dom]["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": float(kavg[dom]),
                "bytes_per_launch": kbytes, "rows_per_launch": rows, "bytes_per_row": b1,
                "kernel_share_of_query": float(kavg[dom] / max(1e-9, float(aa.device_ms)))}
    roofline = kernel_roofline(q, a, man, np.array(ksum[:int(a.launches)]) / args.steps)
    roofline["kernel_share_of_step"] = roofline["kernel_ms"] / ms_per_step
    roofline["timing"] = "CUDA events around the launch on the module's stream, averaged over the %d timed steps" % args.steps
    prof = os.path.join(ROOT, "profiles", "r02_%s_traffic.json" % q)
    if os.path.exists(prof):  # dram__bytes per launch from the committed ncu --set full capture of the same kernel and size
        tr = json.load(open(prof))
        if abs(tr.get("rows", 0) - rows) <= 0.01 * rows:
            roofline["traffic"] = tr.get("dram_bytes_per_launch")
            roofline["traffic_source"] = tr.get("source")
    roofline_q6 = None
    if q == "q1":
        a6, keep6 = mod.prepare("q6", db)
        mod.execute("q6", a6, fetch=True)
        mod.lib.sdqlb200_result_free(ct.byref(a6.result))
        roofline_q6 = kernel_roofline("q6", a6, mod.queries["q6"])
        del a6, keep6
    res_q = mod.run(q, db)
    checked = device_check(q, dcols, res_q, allsum)
    agg_tier = int(a.tier)
    # ---- end to end: plain reference-layout numpy columns in host memory, uploaded + converted every step -----------------
    e2e, cpu_base, host = None, None, None
    if not args.no_e2e:
        try:
            t0 = time.perf_counter()
            host = host_reference_columns(dcols, pin=True)
            t_host = time.perf_counter() - t0
        except Exception as ex:  # noqa: BLE001 -- e.g. the page-locked allocation failed
            e2e = {"value": None, "unit": "GB/s", "error": "host columns: %r" % (ex,)}
    del a, keepalive, db, dcols
    runtime.STORE.clear()
    mod.ws, mod.ws_bytes = None, 0
    torch.cuda.empty_cache()
    if host is not None:
        hdb = reference_db(host)
        runtime.STORE.enabled = False
        fn = getattr(mod, q + "_compiled")
        r = fn(hdb)  # first call: allocator warm-up, workspace
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r = fn(hdb)
        torch.cuda.synchronize()
        e2e_s = allmax((time.perf_counter() - t0) / args.e2e_steps)
        h2d = int(mod.last.h2d_bytes)
        ok = "ok" if (isinstance(r, float) or r.size() == res_q.size()) else "row count differs from the resident run"
        e2e = {"value": all_bytes / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(mod.last.d2h_bytes),
               "host_layout": "numpy int64 / float64 / <U1 (reference layout, 48 B/row in host memory, page-locked); %d B/row cross PCIe" % (h2d // max(1, rows)),
               "result": ok, "host_columns_s": round(t_host, 1),
               "note": "every step: <fn>_compiled(db) with plain numpy columns -> fp64 columns over PCIe as they are while host threads "
                       "narrow the int64 / <U1 columns (int32 / bytes, csrc/sdqlb200_ingest.cu host side), their images follow -> dictionary "
                       "codes on the device -> query -> result rows to the host; %d steps; value = resident-layout bytes / time (same "
                       "numerator as 'value' and as the reference arm)" % args.e2e_steps}
        runtime.STORE.enabled = True
        runtime.STORE.clear()
        # ---- the reference's CPU path on the same host columns (rank 0, N = 1) ------------------------------------------
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                gbs, ms, sample, want = time_reference(q, hdb, rows, bpr, 3, 1, nproc, args.sf)
                from compare import compare
                d = compare(res_q, want)
                cpu_base = {"value": gbs, "unit": "GB/s", "cores": nproc, "kind": "reference", "sample": sample, "ms_per_step": ms,
                            "parity_with_gpu_result": "ok" if d is None else d[:200]}
            except Exception as ex:  # the checker is missing: say so, do not fake a number
                cpu_base = {"value": None, "unit": "GB/s", "cores": nproc, "kind": "reference", "sample": "unavailable: %r" % (ex,)}
        del hdb, host
    # ---- all 22 queries at this scale factor: device latency + parity ----------------------------------------------------
    per_query, fps = {}, load_fingerprints(args.sf)
    if qlist:
        for th in dim_threads:
            th.join()
        t_dims = time.perf_counter() - t_dims
        if dim_err:
            raise dim_err[0][1]
    for qq in qlist:
        mq = mod.queries[qq]
        qdb, held = [], []
        for arg, tname in zip(mq["args"], rr.QUERY_ARGS[qq]):
            if tname in ("lineitem", "orders"):
                cols = dg.columns(tname, needed(mq, arg) + [SCHEMAS[tname][0][0]], orng)
                qdb.append([cols.get(c) for c, _ in SCHEMAS[tname]])
                held.append(cols)
            else:
                qdb.append(dims[tname])
        entry = {}
        try:
            res = mod.run(qq, qdb)
            aq, kq = mod.prepare(qq, qdb)
            ms = []
            for _ in range(args.reps):
                mod.execute(qq, aq, fetch=False)
                ms.append(float(aq.device_ms))
            entry["ms"] = allmax(min(ms))
            entry["launches"] = int(aq.launches)
            entry["rows"] = res.size() if hasattr(res, "size") else 1
            fp = fingerprint(res)
            if qq in fps:
                d = match(fp, fps[qq]["fingerprint"])
                entry["parity"] = ("ok: fingerprint of the reference's result (%s)" % fps[qq].get("source", "oracle/_ref")) if d is None else "MISMATCH: " + d
            else:
                entry["parity"] = "no reference fingerprint committed for SF%g" % args.sf
            if qq in ("q1", "q6") and held:
                dc = device_check(qq, held[0], res, allsum)
                entry["invariants"] = dc
            del aq, kq, res
        except Exception as ex:  # noqa: BLE001 -- one query must not cost the line
            entry["error"] = repr(ex)[:300]
        per_query[qq] = entry
        del qdb, held
        mod.ws, mod.ws_bytes = None, 0  # per-query workspace: the next query may need a very different size
        torch.cuda.empty_cache()
    out = {
        "metric": "tpch_sf%g_%s_scan_throughput" % (args.sf, q), "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sf": args.sf, "query": q, "bytes_per_row": bpr},
        "detail": {"rows_total": rows_total, "rows_this_rank": rows, "layout": bcols, "agg_tier": agg_tier,
                   "partitioning": "lineitem / orders range partitioned on order boundaries over %d rank(s), dimensions replicated" % world,
                   "l2_policy": "inputs (%.2f GB per GPU) larger than L2" % (rows * bpr / 1e9),
                   "merge": None if world == 1 else {"p2p": bool(D.p2p), "merges": int(mod.merges), "p2p_merges": int(mod.p2p_merges),
                                                     "table_merges": int(mod.table_merges)}},
        "roofline": roofline, "roofline_q6": roofline_q6, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "result_check": checked,
        "per_query_ms": {k: round(v["ms"], 4) for k, v in per_query.items() if "ms" in v},
        "per_query": per_query,
        "all_queries_ms": round(sum(v["ms"] for v in per_query.values() if "ms" in v), 3) if per_query else None,
    }
    if qlist:
        out["detail"]["dimension_tables_host_s"] = round(t_dims, 1)
    if cpu_base is not None:
        out["cpu_baseline"] = cpu_base
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
