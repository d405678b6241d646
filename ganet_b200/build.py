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
SOURCES = ["sga.cu", "lga.cu", "volume_ops.cu", "guidance.cu"]

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


PYBIND_SRC = os.path.join(CSRC, "ganet_pybind.cpp")


def pybind_so():
    import sysconfig
    return os.path.join(LIBDIR, "GANet" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pybind(force=False):
    """The reference's native module surface as a real pybind11 module named `GANet` over the C ABI
    (csrc/ganet_pybind.cpp): one host translation unit compiled with g++ against the
    installed torch, linked to libganet_b200.so next to it ($ORIGIN rpath).  torch is only the tensor
    plumbing here; the CUDA library itself has no torch dependency."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension
    build()
    out = pybind_so()
    if not force and not _stale(out, [PYBIND_SRC, SO, os.path.join(HERE, "..", "include", "ganet_b200.h")]):
        return out
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_home = cpp_extension.CUDA_HOME or "/usr/local/cuda"
    inc = cpp_extension.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include")]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-w",
           "-DTORCH_EXTENSION_NAME=GANet", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + p for p in inc]
    cmd += [PYBIND_SRC, "-o", out, "-L" + torch_lib, "-L" + LIBDIR, "-L" + os.path.join(cuda_home, "lib64"),
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lganet_b200",
            "-lcudart", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed on ganet_pybind.cpp")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--pybind", action="store_true", help="also build the pybind11 module `GANet`")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    if a.pybind:
        print(build_pybind(a.force))
