"""host build (tests/emu, -DSDQLB200_EMU) of the device runtime's string primitives: the word-wise shared-memory search
``str_find_w`` must agree with the byte-wise ``str_find`` (= wcsstr semantics of varchar.h:84-97, bounded to the row)."""
import os
import subprocess

from util import ROOT


def test_wordwise_string_search_matches_bytewise(tmp_path):
    exe = os.path.join(tmp_path, "check_strfind")
    src = os.path.join(ROOT, "tests", "emu", "check_strfind.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "sdqlpy_b200", "csrc"), "-I", os.path.join(ROOT, "include"), src, "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout


def test_warp_text_scan_32_lane_path(tmp_path):
    """the 32-lane code path of sdqlrt::warp_text_scan (candidate rows for firstIndex / contains) run on the CPU, one
    std::thread per lane: masks equal the scalar definition bit for bit, no matching row is missed, no byte behind the
    column is read (guard page)."""
    exe = os.path.join(tmp_path, "check_textscan")
    src = os.path.join(ROOT, "tests", "emu", "check_textscan.cpp")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-w", "-I", os.path.join(ROOT, "sdqlpy_b200", "csrc"), src, "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout
