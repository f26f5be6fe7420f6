"""Builds the CPU SIMT emulator of the kernel sources (tests only) into tests/emu/_build/."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libsmfft_emu.so")


def build() -> str:
    srcs = [os.path.join(HERE, "emu_main.cpp"), os.path.join(HERE, "emu_runtime.hpp")]
    srcs += glob.glob(os.path.join(ROOT, "include", "smfft", "detail", "*.cuh"))
    srcs += [os.path.join(ROOT, "smfft_b200", "csrc", f) for f in ("kernels.cuh", "tuning.hpp")]
    if os.path.exists(SO) and all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O1", "-std=c++17", "-DSMFFT_EMU", "-fPIC", "-shared", "-ffp-contract=fast", "-march=x86-64-v3",
           f"-I{ROOT}/include", f"-I{ROOT}/smfft_b200/csrc", f"-I{HERE}", srcs[0], "-o", SO]
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build())
