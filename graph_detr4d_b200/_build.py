"""In-tree build of libgd4d_xview.so (plain nvcc, sm_100a only, no torch headers)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libgd4d_xview.so")
SOURCES = ["c_abi.cu", "xview_fwd.cu", "xview_fwd_tma.cu", "xview_bwd.cu", "xview_bwd_sorted.cu", "xview_v2.cu", "pack.cu", "glue.cu", "gemm_tf32x3.cu", "sgemm_small.cu", "frustum_pe.cu", "match_cost.cu", "fpe.cu"]
HEADERS = ["xview_common.cuh", "xview_records.cuh", "xview_bwd_records.cuh", "xview_sorted_ws.cuh", os.path.join("..", "..", "include", "gd4d_xview.h"),
           os.path.join("..", "..", "include", "gd4d_glue.h"),
           os.path.join("..", "..", "include", "gd4d_frustum.h"),
           os.path.join("..", "..", "include", "gd4d_assign.h"),
           os.path.join("..", "..", "include", "gd4d_fpe.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",            # projection uses explicit _rn intrinsics; bilinear/accumulate may FMA
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libgd4d_xview.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgd4d_xview.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
