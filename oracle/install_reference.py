"""Installs the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so it travels to the GPU box
(``/root/reference`` does not exist there) and ``bench.py``'s CPU arm can time the reference's own modules.

    python -m oracle.install_reference

Step 1 is the stock offline install (``pip install --no-index --no-deps --target baseline/_ref``) from a
scratch copy, because the build writes into the source tree and ``/root/reference`` is read-only.  The
reference's ``setup.py`` uses ``find_packages()`` but its ``models/``, ``utils/`` and ``data_modules/``
directories have no ``__init__.py``, so the wheel holds the top-level modules only; step 2 completes the
install with those three directories, copied verbatim (they import as namespace sub-packages, exactly as
they do from a source checkout).  Nothing under ``baseline/_ref`` is tracked by git or edited.
TEST / BENCH INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("VMTL_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(REPO, "baseline", "_ref")
SUBDIRS = ("models", "utils", "data_modules")


def installed() -> bool:
    return all(os.path.isdir(os.path.join(TARGET, "vision_mtl", d)) for d in SUBDIRS[:2])


def install(force: bool = False) -> str:
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE, "vision_mtl")):
        raise RuntimeError(f"reference tree not found at {REFERENCE}")
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("pip install of the reference failed")
    for d in SUBDIRS:
        dst = os.path.join(TARGET, "vision_mtl", d)
        if not os.path.isdir(dst):
            shutil.copytree(os.path.join(REFERENCE, "vision_mtl", d), dst, ignore=shutil.ignore_patterns("__pycache__"))
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
