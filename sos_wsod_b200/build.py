"""Builds libsoswsod_b200.so (the C-ABI library declared in include/soswsod_b200.h) in-tree with nvcc for
sm_100a.  ``python -m sos_wsod_b200.build`` or ``__graft_entry__.build()``.  The .so is git-ignored but travels
to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsoswsod_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

# (source, extra flags).  The integer/index kernels are built without FMA contraction on top of their explicit
# round-to-nearest intrinsics so that every fp32 comparison they make is IEEE-exact.
SOURCES = [
    ("misc.cu", []),
    ("roi_pool.cu", ["-fmad=false"]),
    ("roi_pool_fast.cu", ["-fmad=false"]),
    ("gemm.cu", []),
    ("wsddn.cu", []),
    ("oicr.cu", ["-fmad=false"]),
    ("nms.cu", ["-fmad=false"]),
    ("tta.cu", ["-fmad=false"]),
    ("pgf.cu", ["-fmad=false"]),
]
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
          "-I", INCLUDE]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsoswsod_b200.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tma.cuh"), os.path.join(CSRC, "roi_plan.cuh"), os.path.join(INCLUDE, "soswsod_b200.h")]

    def compile_one(item):
        src, extra = item
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
