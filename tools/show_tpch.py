#!/usr/bin/env python3
"""print run_tpch.py reports side by side:  python tools/show_tpch.py new.json [old.json]"""
import json
import sys

new = {r["query"]: r for r in json.load(open(sys.argv[1]))}
old = {r["query"]: r for r in json.load(open(sys.argv[2]))} if len(sys.argv) > 2 else {}
tn = to = 0.0
for q, r in new.items():
    ks = {k: v for k, v in r["kernels"].items() if v > 0.08 * r["device_ms_min"]}
    o = old.get(q)
    tn += r["device_ms_min"]
    to += o["device_ms_min"] if o else 0
    bm = r.get("bytes_moved")
    moved = " moved %6.2f GB frac %.3f" % (bm["total"] / 1e9, r["frac_bytes_moved_of_measured_hbm"]) if bm else ""
    print(" %-4s %8.3f ms %s scan %6.2f GB frac %.3f%s ws %7.0f MB rows %-8d %s %s" % (
        q, r["device_ms_min"], ("(was %8.3f)" % o["device_ms_min"]) if o else "", r["scan_bytes"] / 1e9,
        r["frac_of_measured_hbm"], moved, r["workspace_MB"], r["rows"], r.get("parity", ""), ks))
print(" total %.3f ms %s" % (tn, "(was %.3f)" % to if old else ""))
