#!/usr/bin/env python3
"""build an experimental / diagnostic variant of the TPC-H module into gpurun_variants/<tag>.so (git-ignored, travels
to the GPU box):   tools/build_variant.py <tag> [ENV=VALUE ...]
  ENV=VALUE   code-generator switches (SDQLB200_*), set before the generator is imported
  ONLY=q1,q6  queries to include (default q1,q6,q3,q5; ONLY=all = all 22)
  NVCCDEF=A,B preprocessor defines for nvcc, e.g. NVCCDEF=SDQLB200_STATS (the counting build behind the bytes-moved
              roofline, tools/run_tpch.py --stats-so)"""
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
only = os.environ.get("ONLY", "q1,q6,q3,q5")
text, _ = build.compile_source(open(src).read(), "queries.py", only=None if only == "all" else only.split(","))
d = os.path.join(ROOT, "gpurun_variants")
os.makedirs(d, exist_ok=True)
cu = os.path.join(d, tag + ".cu")
open(cu, "w").write(text)
defs = [x for x in os.environ.get("NVCCDEF", "").split(",") if x]
print(build.nvcc(cu, os.path.join(d, tag + ".so"), defines=defs))
