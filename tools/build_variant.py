#!/usr/bin/env python3
"""build an experimental variant of the TPC-H module: tools/build_variant.py <tag> [ENV=VALUE ...] -> <tag>.so"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1]
for kv in sys.argv[2:]:
    k, v = kv.split("=", 1)
    os.environ[k] = v
from sdqlpy_b200 import build  # noqa: E402
src = os.path.join(ROOT, "sdqlpy_b200", "tpch", "queries.py")
text, _ = build.compile_source(open(src).read(), "queries.py", only=os.environ.get("ONLY", "q1,q6,q3,q5").split(","))
d = os.path.join(ROOT, "gpurun_variants")
os.makedirs(d, exist_ok=True)
cu = os.path.join(d, tag + ".cu")
open(cu, "w").write(text)
print(build.nvcc(cu, os.path.join(d, tag + ".so")))
