#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- builds the *real* reference (edin-dal/sdqlpy) TPC-H module.

Recipe.  Nothing from /root/reference is copied into git:
every output of this script lands under ``oracle/_ref/`` which is git-ignored (but travels to the
GPU box with ``gpurun`` like any other built artefact).

  1. scratch copy of /root/reference/src/sdqlpy -> oracle/_ref/site/sdqlpy
     shim 1 (py3.12): ``node.slice.value`` -> ``node.slice``  (sdql_compiler.py:319-332 uses the
     py3.8 ``ast.Index`` wrapper that no longer exists).
  2. workload file = /root/reference/test/test_all.py with (a) the ``sdqlpy_init``/``benchmark``
     driver lines dropped, (b) ``dense(N, ..)`` bounds scaled by SF (test_all.py:185, 397, 461, 1054,
     1065, 1073, 1126 are SF1 key bounds; the reference writes out of bounds above SF1),
     (c) optionally Q15's hard-wired constant (test_all.py:733) replaced.  Deviations (b),(c) are
     the ones BASELINE.md section 2.3 lists.
  3. the reference's own compiler (lib/sdql_compiler.py) generates <name>_compiled.cpp + fast_dict.cpp.
  4. shim 3 (numpy 2): ``PyArray_DATA(`` -> ``PyArray_DATA((PyArrayObject*)``.
  5. g++ -std=c++17 -O3 (flags of fast_dict_generator.py:132-133) against oracle/tbb_shim (shim 2,
     a std::thread stand-in for the six TBB symbols; threads==1 uses plain loops, gen:470-517).

Usage:  python oracle/build_ref.py [--sf 1] [--threads 1] [--q15 <float>] [--opt -O3]
Outputs: oracle/_ref/mods/tpchref_sf<SF>_t<T>_compiled*.so  and  ..._fastdict_compiled*.so
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
SITE = os.path.join(OUT, "site")
MODS = os.path.join(OUT, "mods")

DENSE_SF1 = {6000000: 6000000, 200000: 150000, 150000: 150000}  # bound -> per-SF key range


def modname(sf, threads, tag=""):
    s = ("%g" % sf).replace(".", "p")
    return "tpchref_sf%s_t%d%s" % (s, threads, tag)


def ensure_site():
    dst = os.path.join(SITE, "sdqlpy")
    if not os.path.isdir(dst):
        os.makedirs(SITE, exist_ok=True)
        shutil.copytree(os.path.join(REF, "src", "sdqlpy"), dst)
        subprocess.check_call(["chmod", "-R", "u+w", dst])
        p = os.path.join(dst, "lib", "sdql_compiler.py")
        s = open(p).read().replace("node.slice.value", "node.slice")  # shim 1
        open(p, "w").write(s)
    return dst


def make_workload(sf, q15, name, workdir):
    src = open(os.path.join(REF, "test", "test_all.py")).read()
    out = []
    for line in src.split("\n"):
        if line.startswith("sdqlpy_init(") or line.startswith("benchmark("):
            continue
        out.append(line)
    s = "\n".join(out)
    if sf > 1:
        def repl(m):
            n = int(m.group(1))
            per = DENSE_SF1.get(n, n)
            return "dense(%d," % int(max(n, per * sf))
        s = re.sub(r"dense\((\d+),", repl, s)
    if q15 is not None:
        s = s.replace("max_revenue = 1772627.2087", "max_revenue = %r" % float(q15))
    path = os.path.join(workdir, name + ".py")
    open(path, "w").write(s)
    return path


def build(sf=1, threads=1, q15=None, opt="-O3", tag="", force=False):
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    name = modname(sf, threads, tag)
    os.makedirs(MODS, exist_ok=True)
    so1 = os.path.join(MODS, name + "_compiled" + ext)
    so2 = os.path.join(MODS, name + "_fastdict_compiled" + ext)
    if os.path.exists(so1) and os.path.exists(so2) and not force:
        return name
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s not present; prebuilt %s missing" % (REF, so1))
    pkg = ensure_site()
    work = os.path.join(OUT, "work", name)
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    wl = make_workload(sf, q15, name, work)
    env = dict(os.environ, PYTHONPATH=os.path.join(pkg, "lib"))
    subprocess.check_call([sys.executable, os.path.join(pkg, "lib", "sdql_compiler.py"), os.path.basename(wl), "1",
                           str(threads), pkg + "/"], cwd=work, env=env, stdout=subprocess.DEVNULL)
    cpp = os.path.join(work, name + "_compiled.cpp")
    s = open(cpp).read().replace("PyArray_DATA(", "PyArray_DATA((PyArrayObject*)")  # shim 3
    open(cpp, "w").write(s)
    import numpy
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           "-I" + os.path.join(HERE, "tbb_shim")]
    base = ["g++", "-std=c++17", opt, "-w", "-fPIC", "-shared", "-pthread"] + inc
    p1 = subprocess.Popen(base + [cpp, "-o", so1])
    p2 = subprocess.Popen(base + [os.path.join(work, "fast_dict.cpp"), "-o", so2])
    if p1.wait() or p2.wait():
        raise RuntimeError("g++ failed for " + name)
    return name


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=1)
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--q15", type=float, default=None)
    ap.add_argument("--opt", default="-O3")
    ap.add_argument("--tag", default="")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(build(a.sf, a.threads, a.q15, a.opt, a.tag, a.force))
