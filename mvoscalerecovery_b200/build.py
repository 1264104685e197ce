"""Build libmvosr.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmvosr.so")
SOURCES = ["api.cu"]
HEADERS = ["frame_kernel.cuh", "gstar.cuh", "predicates.cuh", "philox.cuh", "triangulate.cuh", "aux_kernels.cuh",
           "five_point.cuh", "five_point_kernel.cuh", "five_point_tables.h", "bucket_kernel.cuh",
           os.path.join("..", "..", "include", "mvosr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MVOSR_NVCC_EXTRA", "").split()          # e.g. -DMVOSR_STAR_COUNTERS -DMVOSR_DEBUG_PRINT (profiling / debugging builds)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
