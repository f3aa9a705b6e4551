"""TEST-ONLY: build a generated query module as single-threaded host C++ (-DSDQLB200_EMU) and run it through the
normal runtime with a numpy "device" back end.  Purpose: check generated query *logic* against the oracle in the
GPU-less dev container.  Nothing here is imported by the package; the product path raises without CUDA."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_emu(cu_path, so_path, opt="-O1", defines=()):
    cmd = ["g++", "-x", "c++", "-std=c++17", opt, "-w", "-fPIC", "-shared", "-DSDQLB200_EMU"] + ["-D" + d for d in defines] + ["-I", HERE,
           "-I", os.path.join(ROOT, "sdqlpy_b200", "csrc"), "-I", os.path.join(ROOT, "include"), cu_path, "-o", so_path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stderr[-6000:])
    return so_path


class EmuBackend:
    name = "emu"

    def upload(self, arr):
        a = np.ascontiguousarray(arr).copy()
        if a.nbytes == 0:
            a = np.zeros(16, dtype=np.uint8)
        return a.ctypes.data, a

    def upload_packed(self, packed):
        # the emulation "device" is host memory: expand with the numpy restatement of the decode kernels
        a = np.ascontiguousarray(packed.decode_host())
        return a.ctypes.data, a, packed.codes.nbytes

    def alloc(self, nbytes):
        a = np.zeros(max(int(nbytes), 256), dtype=np.uint8)
        return a.ctypes.data, a

    def to_host(self, holder, nbytes):
        return holder.view(np.uint8).reshape(-1)[:int(nbytes)]

    def stream(self):
        return None

    def sync(self):
        pass
