"""In-tree build of libsmfft.so for sm_100a (explicit nvcc; the .so travels to the GPU box with gpurun).

    python -m smfft_b200.build [--force]

One translation unit per FFT size (csrc/inst_e*.cu) plus the host launchers (csrc/launch.cu),
compiled in parallel, linked into smfft_b200/lib/libsmfft.so with the static CUDA runtime.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# experiment builds: SMFFT_VARIANT=name SMFFT_EXTRA_NVFLAGS="-D..." -> lib/libsmfft_name.so (load with SMFFT_LIB=path)
VARIANT = os.environ.get("SMFFT_VARIANT", "")
OBJ = os.path.join(PKG, "lib", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(PKG, "lib", "libsmfft" + ("_" + VARIANT if VARIANT else "") + ".so")
NVCC = os.environ.get("SMFFT_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVFLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    f"-I{os.path.join(ROOT, 'include')}", f"-I{CSRC}",
] + os.environ.get("SMFFT_EXTRA_NVFLAGS", "").split()


def _deps():
    hdr = glob.glob(os.path.join(ROOT, "include", "**", "*"), recursive=True)
    hdr += glob.glob(os.path.join(CSRC, "*.hpp")) + glob.glob(os.path.join(CSRC, "*.cuh"))
    return [h for h in hdr if os.path.isfile(h)]


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src: str, force: bool, extra=()) -> str:
    obj = os.path.join(OBJ, os.path.basename(src).replace(".cu", ".o"))
    if force or _stale(obj, [src] + _deps()):
        cmd = [NVCC, *NVFLAGS, *extra, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit: use the prebuilt library that travelled with the repo
        raise RuntimeError("nvcc not found and no prebuilt libsmfft.so")
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with cf.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"linked {LIB}")
    if not VARIANT:
        _build_compat(force)
    return LIB


def _build_compat(force: bool) -> str:
    """libsmfft_compat.so: the reference's C++ host symbols (include/smfft_compat.hpp) over the C ABI."""
    src = os.path.join(CSRC, "compat", "compat_host.cpp")
    out = os.path.join(PKG, "lib", "libsmfft_compat.so")
    deps = [src, LIB, os.path.join(ROOT, "include", "smfft.h"), os.path.join(ROOT, "include", "smfft_compat.hpp")]
    if force or _stale(out, deps):
        cuda = os.path.dirname(os.path.dirname(NVCC))
        cmd = [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", f"-I{os.path.join(ROOT, 'include')}",
               f"-I{cuda}/include", src, "-o", out, f"-L{os.path.dirname(LIB)}", "-lsmfft", f"-L{cuda}/lib64", "-lcufft", "-lcudart",
               "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{cuda}/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"compat link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
