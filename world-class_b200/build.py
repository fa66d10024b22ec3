"""Builds libworldb200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The .so stays next to this file so that it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libworldb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false",  # FMAs only where spelled out (see wb_common.cuh)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-ccbin", "/usr/bin/g++",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for f in os.listdir(CSRC):
        m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    inc = os.path.join(HERE, "..", "include", "worldb200.h")
    m = max(m, os.path.getmtime(inc), os.path.getmtime(__file__))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(lambda s: _compile(s, verbose), _sources()):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s" % src)
            objs.append(obj)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
