"""Build libamqb.so (the C-ABI CUDA library) in-tree for sm_100a.

    python -m amq_b200.build [--force]

Plain nvcc, no torch headers in any translation unit (seconds per file).  The
.so lands in amq_b200/lib/ (git-ignored, but it travels with the gpurun snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libamqb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
LIBDIR = os.environ.get("AMQB_LIBDIR", LIBDIR)      # tools/build_variant.sh: compile-time variants side by side
LIB = os.path.join(LIBDIR, "libamqb.so")
FLAGS += os.environ.get("AMQB_CFLAGS", "").split()
if os.environ.get("AMQB_TIMELINE") == "1":          # debug build: per-CTA clock stamps in the decode kernel (tools/timeline.py)
    FLAGS.append("-DAMQB_TIMELINE")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "amqb.h"))
    return hdrs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src):
    obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
    if _stale(obj, [src] + _deps()):
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = _sources()
    if force:
        for f in os.listdir(LIBDIR):
            if f.endswith((".o", ".so")):
                os.remove(os.path.join(LIBDIR, f))
    if not _stale(LIB, srcs + _deps()):
        return LIB
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
