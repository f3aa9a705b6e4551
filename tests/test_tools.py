"""the reporting tools run on the committed B200 reports (guards the JSON layout the docs' tables are made from)."""
import os
import subprocess
import sys

from util import ROOT

REPORT = os.path.join(ROOT, "profiles", "r01_tpch_sf100_n1_all22_v6.json")


def run(*args):
    r = subprocess.run([sys.executable] + list(args), capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_roofline_table_and_show():
    out = run("tools/roofline_table.py", REPORT)
    lines = [l for l in out.splitlines() if l.startswith("| q") and l[3].isdigit()]
    assert len(lines) == 22 and out.splitlines()[-1].startswith("| all 22 |")
    assert "| q6 | " in out and "moved frac" in out
    out = run("tools/show_tpch.py", REPORT, os.path.join(ROOT, "profiles", "r01_tpch_sf100_n1_all22_v5.json"))
    assert out.count(" ms ") >= 22 and "total" in out and "moved" in out


def test_tool_scripts_parse():
    import ast
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            ast.parse(open(os.path.join(ROOT, "tools", f)).read(), f)
    for f in ("bench.py", "__graft_entry__.py"):
        ast.parse(open(os.path.join(ROOT, f)).read(), f)
