"""Build libganet_b200.so (the C-ABI CUDA library) in-tree for sm_100a.

    python -m ganet_b200.build [--force] [--verbose]

Plain nvcc, no torch headers: the library's interface is include/ganet_b200.h.
The .so lands in ganet_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
SO = os.path.join(LIBDIR, "libganet_b200.so")
SOURCES = ["sga.cu", "lga.cu", "volume_ops.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps(src):
    yield os.path.join(CSRC, src)
    for f in sorted(os.listdir(CSRC)):          # every header: cheap, and never stale
        if f.endswith((".cuh", ".h")):
            yield os.path.join(CSRC, f)
    yield os.path.join(HERE, "..", "include", "ganet_b200.h")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed on " + src)
    with open(obj + ".ptxas.log", "w") as fh:      # register / spill report per kernel
        fh.write(r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [s for s in SOURCES
            if force or _stale(os.path.join(OBJDIR, s.replace(".cu", ".o")), _deps(s))]
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(todo)) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    if todo or _stale(SO, objs):
        cmd = [_nvcc(), "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
