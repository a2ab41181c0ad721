"""In-tree nvcc build of ``libvmtl_b200.so`` (sm_100a only).

``python -m vision_mtl_b200.build`` or ``__graft_entry__.build()``.  The shared object is a
plain C-ABI library (``include/vmtl_b200.h``): no torch headers, no pybind.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_DIR, "include")
LIB_PATH = os.path.join(PKG_DIR, "libvmtl_b200.so")

SOURCES = ["capi.cu", "xstitch.cu", "metrics.cu", "head_loss.cu", "head_tc.cu", "head_tc_bwd.cu", "gate.cu", "gate_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvmtl_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "vmtl_b200.h")]
    return any(os.path.getmtime(d) > lib_m for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``vision_mtl_b200/libvmtl_b200.so``."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("VMTL_NVCC_EXTRA", "").split()  # e.g. -DVMTL_HT_PROF for the role-timing probes
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    tmp = LIB_PATH + ".tmp"
    cmd += ["-o", tmp, "-lcuda"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libvmtl_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
