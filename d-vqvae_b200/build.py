"""Build libdvq_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python d-vqvae_b200/build.py [--force] [--verbose]

One object per .cu (compiled in parallel), linked into d-vqvae_b200/dvq/libdvq_sm100.so.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# DVQ_BUILD_TAG=<tag> (with DVQ_CFLAGS="-D..." for experiment switches) builds a side-by-side variant
# libdvq_sm100_<tag>.so that `DVQ_LIB=...` selects at import time; the default build has no tag.
TAG = os.environ.get("DVQ_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "dvq", "libdvq_sm100" + ("_" + TAG if TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC] + (["-DDVQ_TC_STATS"] if os.environ.get("DVQ_TC_STATS") else []) + os.environ.get("DVQ_CFLAGS", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "dvq.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return src, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    dep = _deps_mtime()
    todo = []
    for s in srcs:
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(dep, os.path.getmtime(os.path.join(CSRC, s))):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, log in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose:
                    print("==", src)
                    print(log)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
