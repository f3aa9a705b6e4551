"""Algorithmic byte counts of a compiled query.

  scan bytes   every kernel that scans a relation reads each of its streamed columns once, in the resident device
               layout (int32 / fp64 / dictionary codes), plus the fixed-width bytes of the string columns it searches.
  bytes moved  scan bytes + table arrays written by initialisation + one 32-byte sector per data-dependent access
               (slots touched by probes and by insert-or-find, aggregate atomics, late-materialisation gathers) +
               4 bytes per presence-bit test + 8 bytes per result field.  The data-dependent counts come from a
               counting build of the module (-DSDQLB200_STATS, CompiledModule.stats()).

Both are lower bounds on HBM traffic that do not depend on how the kernels are written, so time x peak bandwidth /
bytes is a fair roofline fraction: the scan figure for scan-dominated queries (Q1, Q6), the bytes-moved figure for
the join-heavy ones."""

ELEM_BYTES = {"i32": 4, "f64": 8, "code": 1}
SECTOR = 32


def scan_bytes(man, nrows, code_width=None):
    """-> (column bytes, string bytes) streamed by the relation-scan kernels of one query.
    nrows: {relation argument: rows}; code_width: optional {(arg, column): element bytes} for dictionary codes wider
    than one byte."""
    cols = strs = 0
    for k in man["kernels"]:
        if k["source"][0] != "rel":
            continue
        arg = k["source"][1]
        n = int(nrows[arg])
        for c, rep in k["scan_cols"]:
            w = ELEM_BYTES.get(rep, 1)
            if rep == "code" and code_width:
                w = code_width.get((arg, c), w)
            cols += n * w
        for c, w in k.get("byte_cols", []):
            strs += n * int(w)
    return cols, strs


def bytes_moved(man, nrows, stats, result_rows=0, code_width=None):
    """-> dict with the terms of the bytes-moved figure and their total (see the module docstring)."""
    cols, strs = scan_bytes(man, nrows, code_width)
    random = SECTOR * (stats.get("find_slots", 0) + stats.get("upsert_slots", 0) + stats.get("atomics", 0) +
                       stats.get("gathers", 0))
    out = {"scan_columns": cols, "scan_strings": strs, "tables_initialised": stats.get("init_bytes", 0),
           "random_sectors": random, "presence_bits": 4 * stats.get("bit_tests", 0),
           "result": 8 * int(result_rows) * max(1, len(man.get("result") or []))}
    out["total"] = sum(out.values())
    return out
