"""Builds the CPU SIMT emulator of the kernel sources (tests only) into tests/emu/_build/."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libsmfft_emu.so")


def build() -> str:
    units = [os.path.join(HERE, f) for f in ("emu_main.cpp", "emu_alt.cpp", "emu_dual.cpp")]
    srcs = units + [os.path.join(HERE, "emu_runtime.hpp"), os.path.join(HERE, "emu_run_cfg.hpp")]
    srcs += glob.glob(os.path.join(ROOT, "include", "smfft", "detail", "*.cuh"))
    srcs += [os.path.join(ROOT, "smfft_b200", "csrc", f) for f in ("kernels.cuh", "tuning.hpp")]
    if os.path.exists(SO) and all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    flags = ["-O1", "-std=c++17", "-DSMFFT_EMU", "-fPIC", "-ffp-contract=fast", "-march=x86-64-v3",
             f"-I{ROOT}/include", f"-I{ROOT}/smfft_b200/csrc", f"-I{HERE}"]
    objs = [os.path.join(HERE, "_build", os.path.basename(u).replace(".cpp", ".o")) for u in units]
    procs = [subprocess.Popen([cxx, *flags, "-c", u, "-o", o]) for u, o in zip(units, objs)]  # the units compile in parallel
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("emulator build failed")
    subprocess.run([cxx, "-shared", *objs, "-o", SO], check=True)
    return SO


if __name__ == "__main__":
    print(build())
