"""In-tree build of the CUDA library: `python -m leibnizgym_b200.build`.

nvcc cross-compiles sm_100a without a GPU present.  -fmad=false is part of the
numerics contract (csrc/lg_device.cuh): no implicit FMA contraction, so fp32
operations round exactly as the reference's separate ATen ops do.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", "lg_kernels.cu")]
OUTPUT = os.path.join(HERE, "libleibniz_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"),
    "-I", os.path.join(HERE, "csrc"),
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUTPUT):
        return True
    out_m = os.path.getmtime(OUTPUT)
    csrc = os.path.join(HERE, "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "leibniz_b200.h")]
    return any(os.path.getmtime(d) > out_m for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), output: str = None) -> str:
    """`defines` / `output`: development builds (e.g. -DLG_TRACE into ab/trace.so); the shipped library takes neither."""
    if output is None and not defines and not force and not needs_build():
        return OUTPUT
    out = output or OUTPUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [find_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], *(["-Xptxas", "-v"] if verbose else []), "-o", out, *SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    # python -m leibnizgym_b200.build [--force] [-v] [-DNAME[=V] ...] [-o path]
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outp = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, output=outp))
