"""Compile ``libbdg.so`` (the C-ABI library) for sm_100a with nvcc, in-tree.

``python -m bodge_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU, so this runs in the CPU-only build container; the resulting ``.so`` travels to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbdg.so")
SOURCES = ["assemble.cu", "scan.cu", "cheb.cu", "cheb_ell.cu", "cheb_pair.cu", "observables.cu"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libbdg.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "bdg.h")]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """``defines`` / ``out``: development builds of kernel variants (``-DNAME``) into another file, loaded by
    setting ``BDG_LIB`` (``_native.load``); the product is always ``libbdg.so`` without defines."""
    target = out or LIB
    if not force and not defines and not is_stale():
        return LIB
    cmd = [
        nvcc_path(), "-O3", "-std=c++17", "-lineinfo", "--threads", "0",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC,-O2,-Wall", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-o", target,
    ] + [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
