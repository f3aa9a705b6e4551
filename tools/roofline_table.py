#!/usr/bin/env python3
"""markdown table of a tools/run_tpch.py report (with --stats-so): latency, scan bytes and bytes moved against the
measured HBM peak.   python tools/roofline_table.py profiles/r01_tpch_sf100_n1_all22_v6.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6547.5
rows = {r["query"]: r for r in json.load(open(sys.argv[1]))}
print("| query | device ms | scan GB | scan GB/s | frac of %.0f GB/s | bytes moved GB | moved frac | largest kernels (ms) |" % peak)
print("|---|---|---|---|---|---|---|---|")
tot = 0.0
for i in range(1, 23):
    r = rows.get("q%d" % i)
    if r is None:
        continue
    ms = r["device_ms_min"]
    tot += ms
    bm = r.get("bytes_moved", {}).get("total")
    ks = sorted(r["kernels"].items(), key=lambda kv: -kv[1])[:3]
    print("| q%d | %.2f | %.2f | %.0f | %.2f | %s | %s | %s |" % (
        i, ms, r["scan_bytes"] / 1e9, r["scan_bytes"] / ms / 1e6, r["scan_bytes"] / ms / 1e6 / peak,
        "%.2f" % (bm / 1e9) if bm else "—", "%.2f" % (bm / ms / 1e6 / peak) if bm else "—",
        ", ".join("%s %.2f" % (k.split("_", 1)[1], v) for k, v in ks if v >= 0.05)))
print("| all 22 | %.1f | | | | | | |" % tot)
