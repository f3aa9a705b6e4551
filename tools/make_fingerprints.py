#!/usr/bin/env python3
"""Parity at BASELINE scale factors + the committed fingerprints bench.py checks against (TEST INFRASTRUCTURE: runs the
real reference module oracle/_ref on the box's host cores).

For every query: lineitem / orders are generated in HBM (the generator is bit-identical to the numpy one), copied to the
host in the REFERENCE layout (int64 / float64 / <U n), dimension tables come from the numpy generator; the reference's
generated C++ runs on those columns with all host threads, the CUDA path on the device-resident ones; the two results are
compared in full (tests/compare.py: keys / ints exact, fp64 1e-9) and the fingerprint OF THE REFERENCE's result is written.

   python tools/make_fingerprints.py --sf 100 --queries q1,q6,q3,q5,q9,q18 --out tests/golden/tpch_sf100_fingerprints.json
"""
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


def to_reference_layout(dc, width=None):
    """runtime.DeviceColumn -> numpy array as read_csv would have made it (strings: <U`width`, the schema's VarChar<N> -- the
    reference casts the buffer to VarChar<N>*, sdql_compiler.py:652-668, so a narrower numpy dtype would be misread)"""
    import torch
    n = dc.rows
    if dc.kind == "i32":
        return dc.holder[:n].to(torch.int64).cpu().numpy()
    if dc.kind == "f64":
        return dc.holder[:n].cpu().numpy()
    if dc.kind == "code":
        w = width or max(len(s) for s in dc.dictionary)
        return np.array(dc.dictionary, dtype="<U%d" % max(1, w))[dc.holder[:n].cpu().numpy()]
    m = dc.holder[:n].cpu().numpy()  # fixed-width bytes
    if width and width > m.shape[1]:
        m = np.pad(m, ((0, 0), (0, width - m.shape[1])))
    return np.ascontiguousarray(m.astype(np.uint32)).view("<U%d" % m.shape[1]).reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--queries", default=",".join("q%d" % i for i in range(1, 23)))
    ap.add_argument("--out", required=True)
    ap.add_argument("--report", default=None, help="per-query timing / parity report (profiles/)")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    import torch
    import ref_runner as rr
    from bench import needed
    from compare import compare
    from fingerprint import fingerprint
    from sdqlpy_b200 import runtime
    from sdqlpy_b200.tpch.gen_device import DeviceTPCH
    os.environ["SDQL_REF_THREADS"] = str(a.threads)
    name = next((n for n in (("tpchref_sf100_t8",) if a.sf > 10 else ()) + (("tpchref_sf10_t8",) if a.sf > 1 else ()) + ("tpchref_sf1_t8",)
                 if rr.available(n)), None)
    if name is None or (a.sf > 10 and name != "tpchref_sf100_t8") or (a.sf > 1 and name == "tpchref_sf1_t8"):
        raise SystemExit("no reference module with dense() bounds for SF%g under oracle/_ref (python oracle/build_ref.py --sf .. --threads 8)" % a.sf)
    ref = rr.load(name)
    mod = runtime.load_compiled(os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py"))
    g, dg = TPCH(a.sf), DeviceTPCH(a.sf)
    out = json.load(open(a.out)) if os.path.exists(a.out) else {"sf": a.sf, "queries": {}}
    report, dims = [], {}
    for q in a.queries.split(","):
        man = mod.queries[q]
        ddb, rdb = [], []
        t0 = time.time()
        for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
            need = needed(man, arg)
            if t in ("lineitem", "orders"):
                cols = dg.columns(t, need + [SCHEMAS[t][0][0]])
                ddb.append([cols.get(c) for c, _ in SCHEMAS[t]])
                rel = []
                for c, k in SCHEMAS[t]:
                    if c in cols and (c in need or c == SCHEMAS[t][0][0]):
                        rel.append(to_reference_layout(cols[c], k[1] if isinstance(k, tuple) else None))
                    elif isinstance(k, tuple):
                        rel.append(np.zeros(1, dtype="<U%d" % k[1]))
                    else:
                        rel.append(np.zeros(1, dtype=np.float64 if k == "float" else np.int64))
                rdb.append(rel)
            else:
                key = (t, tuple(need))
                if key not in dims:
                    cc = g.columns(t, need + [SCHEMAS[t][0][0]])
                    dims[key] = ([cc.get(c) for c, _ in SCHEMAS[t]], g.ref_table(t, need))
                ddb.append(dims[key][0])
                rdb.append(dims[key][1])
        t_data = time.time() - t0
        got = mod.run(q, ddb)
        t0 = time.time()
        want = rr.run(ref, q, rdb)
        ref_ms = (time.time() - t0) * 1e3
        d = compare(got, want)
        runs = 1
        if d is not None:  # the threaded reference's dense bool sets race (ref_runner.check): look again
            d, runs = rr.check(ref, q, rdb, got, compare)
            want = rr.run(ref, q, rdb) if d is not None else (got.tuples() if hasattr(got, "tuples") else got)
        row = {"query": q, "sf": a.sf, "parity": "ok" if d is None else d[:200], "ref_ms": round(ref_ms, 1), "ref_threads": a.threads,
               "ref_module": name, "ref_runs": runs, "device_ms": float(mod.last.device_ms), "data_s": round(t_data, 1),
               "rows": got.size() if hasattr(got, "size") else 1}
        print(json.dumps(row), flush=True)
        report.append(row)
        if d is None:
            out["queries"][q] = {"fingerprint": fingerprint(want), "source": "%s, %d TBB-shim threads, %s" % (name, a.threads, time.strftime("%Y-%m-%d")),
                                 "rows": row["rows"]}
        del ddb, rdb, got, want
        runtime.STORE.clear()
        mod.ws, mod.ws_bytes = None, 0
        torch.cuda.empty_cache()
        json.dump(out, open(a.out, "w"), indent=1)
    if a.report:
        json.dump(report, open(a.report, "w"), indent=1)


if __name__ == "__main__":
    main()
