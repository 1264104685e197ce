"""Build libmvosr.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmvosr.so")
# translation units: (source, extra nvcc flags).  The five-point RANSAC is compiled contraction-free (-fmad=false): every FP64
# operation individually rounded, so its results are bit-reproducible on any IEEE-754 host (csrc/five_point_api.cu).
UNITS = [("api.cu", []), ("five_point_api.cu", ["-fmad=false"])]
SOURCES = [u for u, _ in UNITS]
HEADERS = ["handle.h", "frame_kernel.cuh", "gstar.cuh", "gstrip.cuh", "gindex.cuh", "gthread.cuh", "predicates.cuh", "philox.cuh", "triangulate.cuh", "aux_kernels.cuh",
           "five_point.cuh", "five_point_kernel.cuh", "five_point_tables.h", "bucket_kernel.cuh",
           os.path.join("..", "..", "include", "mvosr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


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
    objs = []
    procs = []
    for src, flags in UNITS:                                     # the units compile side by side
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs, cwd=CSRC)
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
