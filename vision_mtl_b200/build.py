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

SOURCES = ["capi.cu", "adam.cu", "bnrelu.cu", "xstitch.cu", "metrics.cu", "head_loss.cu", "head_tc.cu", "head_tc_bwd.cu", "gate.cu", "gate_tc.cu", "upsample.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(PKG_DIR, "build")  # git-ignored; one object per translation unit


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


def _compile_one(nvcc: str, src: str, extra: list, verbose: bool) -> str:
    """One translation unit -> object file; recompiled only when it or a header is newer."""
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps = [path, os.path.join(INCLUDE, "vmtl_b200.h"), os.path.abspath(__file__), *headers]
    flags_tag = obj + ".flags"
    flags = " ".join(NVCC_FLAGS + extra)
    if (os.path.exists(obj) and os.path.exists(flags_tag) and open(flags_tag).read() == flags
            and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps)):
        return obj
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-I", CSRC, "-c", path, "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed on {src}")
    if verbose:
        sys.stderr.write(proc.stderr)
    with open(flags_tag, "w") as f:
        f.write(flags)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (translation units in parallel) and link
    ``vision_mtl_b200/libvmtl_b200.so``."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    extra = os.environ.get("VMTL_NVCC_EXTRA", "").split()  # e.g. -DVMTL_HT_PROF for the role-timing probes
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(lambda s: _compile_one(nvcc, s, extra, verbose), SOURCES))
    tmp = LIB_PATH + ".tmp"
    proc = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", tmp, "-lcuda"],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("linking libvmtl_b200.so failed")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
