"""Builds the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mustache_b200.build            # builds mustache_b200/csrc/libmustache_b200.so
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmustache_b200.so")
SOURCES = ["mb_engine.cu"]
HEADERS = ["mb_kernels.cuh", "mb_normalize.cuh", "mb_sort.cuh", "mb_post.cuh", "mb_parse.h", os.path.join("..", "..", "include", "mustache_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(lib, extra, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed (%d)" % r.returncode)


def build(force=False, verbose=False, lib=LIB, defines=()):
    if force or needs_build(lib):
        _compile(lib, ["-D" + d for d in defines], verbose)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
