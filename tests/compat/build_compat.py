"""Builds tests/compat/_build/libcompat_kernels.so: a user program on include/smfft/compat.cuh (tests only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libcompat_kernels.so")
NVCC = "/usr/local/cuda/bin/nvcc"


def build() -> str:
    src = os.path.join(HERE, "compat_kernels.cu")
    import glob

    deps = [src] + glob.glob(os.path.join(ROOT, "include", "smfft", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "smfft", "detail", "*.cuh"))
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    if not os.path.exists(NVCC):
        if os.path.exists(SO):
            return SO
        raise RuntimeError("nvcc not found and no prebuilt libcompat_kernels.so")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run([NVCC, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-ccbin", "/usr/bin/g++",
                    "-Xcompiler", "-fPIC", "-shared", f"-I{ROOT}/include", src, "-o", SO], check=True)
    return SO


if __name__ == "__main__":
    print(build())
