"""device-side `.tbl` reader (csrc/sdqlb200_tbl.cu) against the reference's read_csv.

CPU: the CUDA source is built as single-threaded host C++ (tests/emu, -DSDQLB200_EMU) and driven through the same host
code (sdqlpy_b200/tbl.py) with the numpy "device" back end; expected values are (a) tests/golden/tbl_fixture.json, made
by the REAL reference read_csv (tests/golden/make_tbl_golden.py), and (b) this package's mirror of it
(sdql_lib.read_csv) on generated tables.  GPU (-m gpu): the same comparisons through libsdqlb200_tbl.so on cuda:0."""
import json
import os
import subprocess

import numpy as np
import pytest

import emu
from sdqlpy_b200 import runtime, sdql_lib, tbl
from sdqlpy_b200.tpch.gen import SCHEMAS, TPCH, bytes_to_ustr
from util import ROOT

FIXTURE = os.path.join(ROOT, "tests", "golden", "tbl_fixture.json")


@pytest.fixture(scope="module")
def emu_tbl(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu_tbl")
    so = os.path.join(d, "tbl_emu.so")
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-w", "-fPIC", "-shared", "-DSDQLB200_EMU", "-I", os.path.join(ROOT, "tests", "emu"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "sdqlpy_b200", "csrc", "sdqlb200_tbl.cu"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    old = runtime._backend
    be = emu.EmuBackend()
    runtime.set_backend(be)
    runtime.STORE.clear()
    yield tbl.lib(so), be
    runtime.set_backend(old)
    runtime.STORE.clear()


def check_against_fixture(library, be, tmp_path, block_bytes=1 << 30):
    fx = json.load(open(FIXTURE))
    for table, t in fx["tables"].items():
        path = os.path.join(tmp_path, table + ".tbl")
        open(path, "w", newline="\n").write(t["text"])
        cols = tbl.parse_file(path, SCHEMAS[table], None, "|", block_bytes, be, library)
        for name, kind in SCHEMAS[table]:
            want, got = t["columns"][name], cols[name]
            if isinstance(kind, tuple):
                assert got.kind == "bytes" and got.data.shape == (len(want), kind[1])
                assert [str(x) for x in bytes_to_ustr(got.data, kind[1])] == want, (table, name)
            elif kind == "float":
                assert got.kind == "f64"
                assert [float(x).hex() for x in got.data] == want, (table, name)   # bit exact
            else:
                assert got.kind == "i32" and [int(x) for x in got.data] == want, (table, name)


def test_golden_fixture_of_the_reference_reader(emu_tbl, tmp_path):
    check_against_fixture(emu_tbl[0], emu_tbl[1], tmp_path)


def test_golden_fixture_in_small_blocks(emu_tbl, tmp_path):
    """the file is cut into blocks of whole rows (here ~1.5 KB each); the pieces are concatenated"""
    check_against_fixture(emu_tbl[0], emu_tbl[1], tmp_path, block_bytes=1500)


@pytest.mark.parametrize("table", ["lineitem", "orders", "part", "supplier", "nation"])
def test_generated_tables_match_read_csv(emu_tbl, tmp_path, table):
    L, be = emu_tbl
    g = TPCH(0.002)
    schema = SCHEMAS[table]
    ref_cols = g.ref_table(table, [c for c, _ in schema])
    path = os.path.join(tmp_path, table + ".tbl")
    open(path, "wb").write(tbl.format_tbl(schema, ref_cols))
    rec = sdql_lib.record({c: (sdql_lib.string(k[1]) if isinstance(k, tuple) else {"int": int, "float": float, "date": sdql_lib.date}[k])
                           for c, k in schema})
    want = sdql_lib.read_csv(path, {rec: bool}, table).getContainer()["data"]
    got = tbl.parse_file(path, schema, None, "|", 1 << 30, be, L)
    for (name, kind), w in zip(schema, want):
        c = got[name]
        if isinstance(kind, tuple):
            assert (bytes_to_ustr(c.data, kind[1]) == w).all(), name
        elif kind == "float":
            assert c.data.dtype == np.float64 and (c.data.view(np.int64) == w.view(np.int64)).all(), name
        else:
            assert c.data.dtype == np.int32 and (c.data == w).all(), name


def parse_bytes(emu_tbl, text, schema, want=None):
    L, be = emu_tbl
    return tbl.parse_text(np.frombuffer(text, dtype=np.uint8), schema, want, "|", be, L)


def test_edge_cases(emu_tbl):
    schema = [("a", "int"), ("b", "float"), ("c", "date"), ("d", ("str", 5)), ("x_NA", ("str", 1))]
    n, host, dev = parse_bytes(emu_tbl, b"", schema)
    assert n == 0 and host["a"][1].shape == (0,)
    # last row without a newline; CRLF; negative values; skipped column; string cut to the width / zero padded
    text = b"1|2.50|1995-06-17|abcdefgh|\r\n-7|-0.01|1992-01-01||\n2147483647|100|1998-12-31|xy|"
    n, host, dev = parse_bytes(emu_tbl, text, schema, want={"a", "b", "c", "d"})
    assert n == 3
    assert host["a"][1].tolist() == [1, -7, 2147483647]
    assert host["b"][1].tolist() == [2.5, -0.01, 100.0]
    assert host["c"][1].tolist() == [19950617, 19920101, 19981231]
    assert [bytes(r) for r in host["d"][1]] == [b"abcde", b"\0\0\0\0\0", b"xy\0\0\0"]
    assert (dev["a"].min, dev["a"].max) == (-7, 2147483647) and (dev["c"].min, dev["c"].max) == (19920101, 19981231)
    assert "x_NA" not in host
    # many rows: row offsets across tile boundaries (4 KB tiles) and across chunk boundaries inside a tile
    rows = [b"%d|%d.%02d|1994-01-%02d|r%d|" % (i, i * 7, i % 100, 1 + i % 28, i % 1000) for i in range(5000)]
    n, host, dev = parse_bytes(emu_tbl, b"\n".join(rows) + b"\n", schema)
    assert n == 5000 and host["a"][1].tolist() == list(range(5000))
    assert host["b"][1].tolist() == [float("%d.%02d" % (i * 7, i % 100)) for i in range(5000)]
    assert host["c"][1].tolist() == [19940100 + 1 + i % 28 for i in range(5000)]


@pytest.mark.parametrize("text,msg", [
    (b"1|2.5|1995-06-17|\n", "fewer fields"),           # the string column is missing
    (b"1|2.5|1995-06-17|abc|\n\n2|1.0|1995-06-17|abc|\n", "row 1"),  # empty line in the middle
    (b"x1|2.5|1995-06-17|abc|\n", "not a plain decimal"),
    (b"1|2.5e3|1995-06-17|abc|\n", "not a plain decimal"),
    (b"1|2.5|1995-06-17|abc|\n3000000000|1|1995-06-17|abc|\n", "row 1"),
    (b"1|1234567890.1234567|1995-06-17|abc|\n", "significant digits"),
    (b"1|2.5|1995-06-17|caf\xc3\xa9|\n", "0x80"),
])
def test_malformed_rows_raise_like_the_reference_would(emu_tbl, text, msg):
    schema = [("a", "int"), ("b", "float"), ("c", "date"), ("d", ("str", 5)), ("x_NA", ("str", 1))]
    with pytest.raises(ValueError) as e:
        parse_bytes(emu_tbl, text, schema)
    assert msg in str(e.value)


def test_read_tbl_feeds_queries(emu_tbl, tmp_path, monkeypatch):
    """read_tbl -> columnar sr_dict whose columns run through a compiled query (emulation build) like generated ones"""
    from sdqlpy_b200 import build
    from util import QUERY_SCRIPT, golden
    from compare import compare
    L, be = emu_tbl
    monkeypatch.setattr(tbl, "_lib", L)
    g = TPCH(0.01)
    schema = SCHEMAS["lineitem"]
    path = os.path.join(tmp_path, "lineitem.tbl")
    open(path, "wb").write(tbl.format_tbl(schema, g.ref_table("lineitem", [c for c, _ in schema])))
    rec = sdql_lib.record({c: (sdql_lib.string(k[1]) if isinstance(k, tuple) else {"int": int, "float": float, "date": sdql_lib.date}[k])
                           for c, k in schema})
    li = sdql_lib.read_tbl(path, {rec: bool}, "li")
    data = li.getContainer()["data"]
    text, _ = build.compile_source(open(QUERY_SCRIPT).read(), "queries.py", only=["q1", "q6"])
    cu = os.path.join(tmp_path, "q.cu")
    open(cu, "w").write(text)
    mod = runtime.CompiledModule(emu.build_emu(cu, os.path.join(tmp_path, "q.so")))
    h2d0 = runtime.STORE.h2d_bytes
    got6 = mod.run("q6", [data])
    got1 = mod.run("q1", [data])
    assert compare(got6, golden(0.01)["q6"]) is None
    assert compare(got1, golden(0.01)["q1"]) is None
    # numeric columns were already device resident after the parse; only the flag columns (bytes -> dictionary codes) went up
    assert runtime.STORE.h2d_bytes - h2d0 <= 2 * len(data[0].data)


@pytest.mark.gpu
def test_gpu_reader_matches_reference_fixture_and_read_csv(tmp_path):
    import torch
    assert torch.cuda.is_available()
    runtime.set_backend(None)
    runtime.STORE.clear()
    L, be = tbl.lib(), runtime.backend()
    check_against_fixture(L, be, tmp_path)
    check_against_fixture(L, be, tmp_path, block_bytes=1500)
    g = TPCH(0.05)
    for table in ("lineitem", "orders", "customer"):
        schema = SCHEMAS[table]
        ref_cols = g.ref_table(table, [c for c, _ in schema])
        path = os.path.join(tmp_path, table + "_big.tbl")
        open(path, "wb").write(tbl.format_tbl(schema, ref_cols))
        got = tbl.parse_file(path, schema, None, "|", 8 << 20, be, L)   # several blocks
        for (name, kind), w in zip(schema, ref_cols):
            c = got[name]
            if isinstance(kind, tuple):
                assert (bytes_to_ustr(c.data, kind[1]) == w).all(), (table, name)
            elif kind == "float":
                assert (c.data.view(np.int64) == np.asarray(w, dtype=np.float64).view(np.int64)).all(), (table, name)
            else:
                assert (c.data == w).all(), (table, name)
    with pytest.raises(ValueError):
        tbl.parse_text(np.frombuffer(b"1|x|\n", dtype=np.uint8), [("a", "int"), ("b", "float")], None, "|", be, L)


def test_library_exports_declared_symbols():
    """the sm_100a build of the reader loads on the CPU and exports every symbol include/sdqlb200_tbl.h declares"""
    import ctypes
    import re
    from sdqlpy_b200 import build
    so = build.compile_tbl()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "sdqlb200_tbl.h")).read()
    names = set(re.findall(r"\b(sdqlb200_tbl_[a-z_0-9]+)\s*\(", hdr))
    assert names >= {"sdqlb200_tbl_scratch_bytes", "sdqlb200_tbl_index", "sdqlb200_tbl_row_starts", "sdqlb200_tbl_parse",
                     "sdqlb200_tbl_last_error"}
    for n in names:
        assert hasattr(lib, n), n
    lib.sdqlb200_tbl_scratch_bytes.restype = ctypes.c_int64
    lib.sdqlb200_tbl_scratch_bytes.argtypes = [ctypes.c_int64]
    assert lib.sdqlb200_tbl_scratch_bytes(0) >= 256 and lib.sdqlb200_tbl_scratch_bytes(1 << 30) >= (1 << 30) // 4096 * 12
    assert ctypes.sizeof(tbl.TblStatus) == 8 * (2 + 2 * tbl.MAX_COLS) and ctypes.sizeof(tbl.TblCol) == 16


def test_blank_lines_are_skipped_like_csv_reader_does(emu_tbl, tmp_path):
    """csv.reader yields [] for a blank line and the reference's row loop adds nothing for it (sdql_lib.py:79-82): the same
    columns with blank lines sprinkled over the text, in one block and in small blocks (a blank line at a block boundary)"""
    import random
    fx = json.load(open(FIXTURE))
    table = "lineitem" if "lineitem" in fx["tables"] else sorted(fx["tables"])[0]
    t = fx["tables"][table]
    rows = t["text"].split("\n")
    assert rows[-1] == ""
    rnd = random.Random(5)
    noisy = ["", ""]
    for r in rows[:-1]:
        noisy.append(r)
        if rnd.random() < 0.3:
            noisy += [""] * rnd.randint(1, 3)
    path = os.path.join(tmp_path, "blank.tbl")
    open(path, "w", newline="\n").write("\n".join(noisy) + "\n\n")
    for block in (1 << 30, 700):
        cols = tbl.parse_file(path, SCHEMAS[table], None, "|", block, emu_tbl[1], emu_tbl[0])
        for name, kind in SCHEMAS[table]:
            want, got = t["columns"][name], cols[name]
            if isinstance(kind, tuple):
                assert [str(x) for x in bytes_to_ustr(got.data, kind[1])] == want, name
            elif kind == "float":
                assert [float(x).hex() for x in got.data] == want, name
            else:
                assert [int(x) for x in got.data] == want, name
