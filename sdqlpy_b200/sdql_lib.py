"""User-facing API -- same names and call conventions as the reference's ``sdqlpy.sdql_lib``
(/root/reference/src/sdqlpy/sdql_lib.py): ``sdqlpy_init`` (lib:372), ``sdql_compile`` (lib:389), ``read_csv``
(lib:118), ``sr_dict`` (lib:132), ``record`` (lib:268), ``vector`` (lib:298), ``string``/``date`` (lib:14-21), the
helper functions (lib:341-368) and ``benchmark`` (lib:437).

What changes underneath: execution modes 1 and 2 dispatch to the B200 backend (IR -> sm_100a CUDA, see
codegen.py / runtime.py) instead of the TBB/phmap C++ module.  Mode 0 stays a pure-Python interpreter of the
DSL and is only meant for tiny inputs.
"""
import csv
import functools
import inspect
import os
import statistics
import sys
import time

import numpy as np


class string:
    def __init__(self, max_size=25):
        self.max_size = max_size


class date:
    pass


# --------------------------------------------------------------------------------------------
# semi-ring dictionary / record / vector (Python-mode semantics, lib:132-337)
# --------------------------------------------------------------------------------------------
class sr_dict:
    def __init__(self, initializer_dict=None, value=None, columnar_layout=False):
        if value is None:
            self._c = {} if initializer_dict is None else initializer_dict
        else:
            self._c = {initializer_dict: value}
        self._columnar = columnar_layout

    def getContainer(self):
        return self._c

    def getColumnarLayoutStatus(self):
        return self._columnar

    def get(self, key):
        return self._c.get(key)

    __getitem__ = get

    def __len__(self):
        return len(self._c)

    def __iter__(self):
        return iter(self._c)

    def items(self):
        return self._c.items()

    def __str__(self):
        def q(x):
            return '"%s"' % x if isinstance(x, (str, np.str_)) else str(x)
        return "{ " + ", ".join("%s: %s" % (q(k), q(v)) for k, v in self._c.items()) + " }"

    __repr__ = __str__

    def __hash__(self):
        return hash(str(self))

    def __eq__(self, other):
        if other is None:
            return self._c is None
        return isinstance(other, sr_dict) and self._c == other._c

    def __add__(self, other):  # key-wise addition (lib:186-198)
        if len(self._c) == 0:
            return other
        if len(other._c) == 0:
            return self
        for k, v in other._c.items():
            if k in self._c:
                self._c[k] = self._c[k] + v
            else:
                self._c[k] = v
        return self

    def _rows(self):
        if not self._columnar:
            yield from self._c.items()
        else:
            heads, data = self._c["headers"], self._c["data"]
            for i in range(len(data[0])):
                yield record({h: _py(data[j][i]) for j, h in enumerate(heads)}), True

    def sum(self, func, is_an_update_sum=True):
        result = None
        for k, v in self._rows():
            r = func((k, v))
            if isinstance(r, dict):
                r = sr_dict(r)
            if r is None or (isinstance(r, sr_dict) and len(r) == 0 and result is not None):
                continue
            result = r if result is None else result + r
        return result

    def joinBuild(self, col, filter, outCols):
        def build(rec):
            out = {c: rec[c] for c in (outCols or [col])}
            return {rec[col]: record(out)}
        return self.sum(lambda p: build(p[0]) if filter(p) else sr_dict())

    def joinProbe(self, indexedDict, col, filter, outputFunc, is_an_update_sum=True):
        def inner(p):
            v = indexedDict[p[0][col]]
            return outputFunc(v, p[0]) if v is not None else sr_dict()
        return self.sum(lambda p: inner(p) if filter(p) else sr_dict(), is_an_update_sum)


def _py(v):
    if isinstance(v, np.generic):
        return v.item()
    return v


class record(sr_dict):
    def __init__(self, initializer_dict=None):
        sr_dict.__init__(self, initializer_dict if initializer_dict is not None else {})

    def __getattr__(self, attr):
        if attr.startswith("_"):
            raise AttributeError(attr)
        return self._c.get(attr)

    def __eq__(self, other):
        if other is None:
            return False
        return list(self._c.values()) == list(other._c.values())

    def __hash__(self):
        return hash(tuple(self._c.values()))

    def __add__(self, other):
        return record({k: v + other._c[k] for k, v in self._c.items()})

    def concat(self, other):
        d = dict(self._c)
        d.update(other._c)
        return record(d)


class vector:
    def __init__(self, initializer_list=None):
        self._c = list(initializer_list) if initializer_list is not None else []

    def getContainer(self):
        return self._c

    def __len__(self):
        return len(self._c)

    def __add__(self, other):
        return vector(self._c + other._c)

    def __str__(self):
        return "[ " + ", ".join(str(v) for v in self._c) + " ]"


def extractYear(d):
    return d // 10000


def firstIndex(s, kw):
    return s.find(kw)


def startsWith(s, kw):
    return s.startswith(kw)


def endsWith(s, kw):
    return s.endswith(kw)


def dictSize(d):
    return len(d.getContainer()) if isinstance(d, (sr_dict, vector)) else len(d)


def substr(s, start, end):
    return s[start:end + 1]


def unique(x):
    return x


def dense(n, x):
    return x


# --------------------------------------------------------------------------------------------
# loading (lib:69-128): pipe-delimited .tbl -> columnar sr_dict({"headers", "data"})
# --------------------------------------------------------------------------------------------
def read_csv(file_path, header_type_dict, dataset_name, delimiter="|"):
    rec_type = list(header_type_dict.keys())[0].getContainer()
    heads, types = list(rec_type.keys()), list(rec_type.values())
    cols = [[] for _ in heads]
    with open(file_path, newline="\n") as f:
        for row in csv.reader(f, delimiter=delimiter):
            for i in range(len(heads)):
                v = row[i] if i < len(row) else ""
                if types[i] is date:
                    cols[i].append(int(v.replace("-", "")))
                elif isinstance(types[i], string):
                    cols[i].append(v)
                else:
                    cols[i].append(types[i](v))
    data = []
    for i, t in enumerate(types):
        if isinstance(t, string):
            data.append(np.array(cols[i], dtype="<U%d" % t.max_size))
        elif t is float:
            data.append(np.array(cols[i], dtype=np.float64))
        else:
            data.append(np.array(cols[i], dtype=np.int64))
    return sr_dict({"headers": heads, "data": data}, None, True)


def read_tbl(file_path, header_type_dict, dataset_name, delimiter="|", columns=None):
    """``read_csv`` with the parsing done on the GPU (sdqlpy_b200/tbl.py, csrc/sdqlb200_tbl.cu): same arguments, same
    columnar ``sr_dict`` out, but the columns are compact (int32 / fp64 / fixed-width bytes) ``Column`` objects whose
    device copies already sit in the column store.  ``columns``: optional subset of column names to parse (the others
    stay ``None``; the first column is always parsed, it carries the row count)."""
    from . import tbl
    rec_type = list(header_type_dict.keys())[0].getContainer()
    heads, types = list(rec_type.keys()), list(rec_type.values())
    schema = []
    for h, t in zip(heads, types):
        if isinstance(t, string):
            schema.append((h, ("str", t.max_size)))
        elif t is date:
            schema.append((h, "date"))
        elif t is float:
            schema.append((h, "float"))
        elif t in (int, bool):
            schema.append((h, "int"))
        else:
            raise ValueError("column %s: unsupported type %r" % (h, t))
    want = None if columns is None else set(columns) | {heads[0]}
    cols = tbl.parse_file(file_path, schema, want, delimiter)
    return sr_dict({"headers": heads, "data": [cols.get(h) for h in heads]}, None, True)


def table_from_columns(headers, data):
    """columnar sr_dict from ready-made columns (numpy arrays and/or device-resident column handles)."""
    return sr_dict({"headers": list(headers), "data": list(data)}, None, True)


# --------------------------------------------------------------------------------------------
# init / dispatch (lib:372-435)
# --------------------------------------------------------------------------------------------
_state = {"mode": 0, "gpus": 1, "engine": None, "partitioned": ("li", "ord"), "partkeys": ("l_orderkey", "o_orderkey")}


def sdqlpy_init(execution_mode=0, threads_count=1):
    """0: run in Python | 1: compile (IR -> CUDA -> nvcc) and run on B200 | 2: run the previously compiled module.
    ``threads_count`` (TBB threads in the reference, fixed for the run, lib:372-387) is the number of GPUs here: with
    N > 1 the queries run on N GPUs of this process (runtime.Engine: one host thread per GPU, relations range
    partitioned, partial dictionaries merged over NVLink)."""
    if execution_mode not in (0, 1, 2):
        print("Execution mode is not supported. Failed.")
        return
    if _state["engine"] is not None and _state["engine"].n != threads_count:
        _state["engine"].close()
        _state["engine"] = None
    _state["mode"], _state["gpus"] = execution_mode, int(threads_count)
    if execution_mode == 1:
        caller = inspect.stack()[1][0].f_code.co_filename
        from . import build
        build.compile_file(caller, force=True)


def set_partitioning(partitioned=("li", "ord"), partkeys=("l_orderkey", "o_orderkey")):
    """extension (no reference counterpart): which relation arguments the multi-GPU engine range partitions and on which
    columns; ``partitioned=None`` = per query the argument with the most rows.  Default: the TPC-H workload's names."""
    _state["partitioned"], _state["partkeys"] = partitioned, tuple(partkeys)
    if _state["engine"] is not None:
        _state["engine"].close()
        _state["engine"] = None


def _engine():
    if _state["gpus"] > 1 and _state["engine"] is None:
        from . import runtime
        _state["engine"] = runtime.Engine(_state["gpus"], _state["partitioned"], _state["partkeys"])
    return _state["engine"] if _state["gpus"] > 1 else None


def sdql_compile(in_type):
    def deco(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if _state["mode"] == 0:
                return func(*args, **kwargs)
            from . import runtime
            mod = runtime.load_compiled(inspect.getfile(func))
            fname = func.__name__ + "_compiled"
            if not hasattr(mod, fname):
                print("Error: the compiled version of " + func.__name__ + " is not found!")
                return None
            db = [a["data"] for a in args]  # columnar layout (lib:420-424)
            eng = _engine()
            if eng is not None:
                return eng.run(mod, func.__name__, db)
            return getattr(mod, fname)(db)
        wrapper.__sdql_in_type__ = in_type
        wrapper.__sdql_func__ = func
        return wrapper
    return deco


def benchmark(title, iterations, func, args, show_results=True, verbose=True):
    """1 untimed warm-up, ``iterations`` timed end-to-end calls (wall clock, ms), one more call for the result
    (protocol of lib:445-452; the SMT toggle of lib:438-442 has no GPU meaning and is dropped)."""
    func(*args)
    times = []
    for _ in range(iterations):
        t0 = time.time() * 1000
        func(*args)
        times.append(time.time() * 1000 - t0)
    res = func(*args)
    mean = sum(times) / max(1, len(times))
    if verbose:
        sd = statistics.stdev(times) if len(times) > 1 else 0.0
        print(title + ": Mean: %0.2f | StDev: %0.2f" % (mean, sd))
        if show_results:
            print(res)
        try:
            print("Result Size: " + str(res.size() if hasattr(res, "size") else len(res)))
        except TypeError:
            print("Scalar Result")
        print("=" * 76)
    else:
        print(title + "\t%0.2f" % mean)
    return times
