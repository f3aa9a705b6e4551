#!/usr/bin/env python3
"""condense an .ncu-rep (ncu --set full) into the few numbers the roofline discussion uses.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]
STALL = "smsp__pcsamp_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name)
    for k in KEEP:
        if k in hdr:
            print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(h[len(STALL):], float(r[i])) for i, h in enumerate(hdr) if h.startswith(STALL) and "not_issued" not in h and r[i]]
    tot = sum(v for _, v in st) or 1
    print("  warp stall samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for n, v in sorted(st, key=lambda x: -x[1])[:7]))
