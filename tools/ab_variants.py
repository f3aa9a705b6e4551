#!/usr/bin/env python3
"""A/B of experimental builds of the TPC-H module on cuda:0: data is generated once per query, every variant
(gpurun_variants/<tag>.so, 'default' = the in-tree build) runs on the same device-resident columns.
   python tools/ab_variants.py --sf 10 --queries q1,q6 --variants default,legacy,ring4 --out gpurun_out/ab.json"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--queries", default="q1,q6,q3,q5,q9,q18")
    ap.add_argument("--variants", default="default,legacy")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default=None)
    ap.add_argument("--device-gen", action="store_true", help="lineitem / orders generated on the GPU (as tools/run_tpch.py): "
                    "no host generation of the fact tables, SF100 fits one GPU")
    a = ap.parse_args()
    import ref_runner as rr
    from compare import compare
    from sdqlpy_b200 import build, runtime
    mods = {}
    for v in a.variants.split(","):
        so = build.out_paths(os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py"))[1] if v == "default" else \
            os.path.join(ROOT, "gpurun_variants", v + ".so")
        mods[v] = runtime.CompiledModule(so)
    g = TPCH(a.sf)
    dg = None
    if a.device_gen:
        from sdqlpy_b200.tpch.gen_device import DeviceTPCH
        dg = DeviceTPCH(a.sf)
    tabs, report = {}, []
    for q in a.queries.split(","):
        base = None
        for v, mod in mods.items():
            if q not in mod.queries:
                continue
            man = mod.queries[q]
            db = []
            for arg, t in zip(man["args"], rr.QUERY_ARGS[q]):
                need = sorted({c for aa, c, r in man["inputs"] if aa == arg} |
                              {x.split(":")[3] for _, x in man["result"] if x.startswith("str:") and x.split(":")[2] == arg and len(x.split(":")) > 3})
                key = (t, tuple(need))
                if key not in tabs:
                    src = dg if (dg is not None and t in ("lineitem", "orders")) else g
                    cols = src.columns(t, need + [SCHEMAS[t][0][0]])
                    tabs[key] = [cols.get(c) for c, _ in SCHEMAS[t]]
                db.append(tabs[key])
            res = mod.run(q, db)
            if base is None:
                base = res
                par = "base"
            else:
                d = compare(res, base.tuples() if hasattr(base, "tuples") else base)
                par = "ok" if d is None else d[:160]
            args_, keep = mod.prepare(q, db)
            ms, kms = [], []
            for _ in range(a.reps):
                mod.execute(q, args_, fetch=False, kernel_times=True)
                ms.append(float(args_.device_ms))
                kms.append([args_.kernel_ms[k] for k in range(int(args_.launches))])
            kmin = np.min(np.array(kms), axis=0)
            row = {"query": q, "variant": v, "sf": a.sf, "device_ms_min": min(ms), "device_ms_mean": float(np.mean(ms)),
                   "vs_first_variant": par,
                   "kernels": {man["kernels"][k]["name"]: round(float(kmin[k]), 4) for k in range(len(kmin))}}
            print(json.dumps(row), flush=True)
            report.append(row)
        tabs.clear()
        runtime.STORE.clear()
        if dg is not None:  # per-query workspaces and fact tables: the next query may need a very different size
            for mod in mods.values():
                mod.ws, mod.ws_bytes = None, 0
            import torch
            torch.cuda.empty_cache()
    if a.out:
        json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
