"""Build the in-tree CUDA library for sm_100a (cross-compiles without a GPU).

    python -m tacotron2_subword_b200.build [--force]

Output: tacotron2_subword_b200/csrc/libtaco2dec.so (git-ignored; travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libtaco2dec.so")
SOURCES = [os.path.join(CSRC, "taco2dec.cu")]
HEADERS = [os.path.join(ROOT, "include", "taco2dec.h")] + sorted(
    os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def up_to_date() -> bool:
    if not os.path.isfile(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def source_stamp() -> str:
    """sha256 (16 hex digits) over the kernel sources the library is built from.  nvcc output is not byte-reproducible, so
    profile captures are tied to the sources (bench.py accepts a capture only if this stamp matches and the in-tree library
    is not older than the sources)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), *SOURCES, "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(CSRC, "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
