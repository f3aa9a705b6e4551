#!/usr/bin/env python3
"""hottest SASS instructions of an `ncu --page source --csv` export:  python tools/ncu_hot.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body) or 1
execd = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("instructions: %d static, %d executed (warp level); %d samples" % (len(body), execd, tot))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
top = sorted(range(len(body)), key=lambda k: -int(body[k][ix["# Samples"]] or 0))[:N]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for k in sorted(top):
    r = body[k]
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5d %5.1f%% exec %9s thr %4s  %-60s %s" % (k, 100.0 * s / tot, r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]],
                                                   r[ix["Source"]].strip()[:60], ", ".join("%s %d" % (n, v) for v, n in why if v)))
