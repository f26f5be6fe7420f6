#!/usr/bin/env python
"""tools/sass_report.py -- static SASS evidence per product kernel instance, as markdown (runs here, no GPU).

    python tools/sass_report.py > profiles/r02_sass_stats.md

For libsmfft.so (forward, table twiddles) and the reference-contract wrapper kernels (tests/compat/_build): instruction
totals per thread and the mnemonics that prove the design: UTMALDG / UTMASTG (TMA tensor copies), SYNCS (mbarrier),
FADD2 / FFMA2 (packed f32x2), LDS / STS (shared-memory exchanges), BAR, SHFL (warp shuffles), MUFU, LDG / STG."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTMALDG", "UTMASTG", "SYNCS", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "FSEL", "MUFU", "FADD", "FADD2", "FMUL", "FFMA", "FFMA2"]
MODES = {0: "C2C", 1: "R2C", 2: "C2R"}
IOS = {0: "tma", 1: "ldg", 2: "tma_stg", 3: "reg"}


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, c = None, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, c
            name, c = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            c["total"] += 1
            c[m.group(1)] += 1
    if name:
        yield name, c


def row(label, c):
    return f"| {label} | {c['total']} | " + " | ".join(str(c[k]) for k in KEYS) + " |"


def main():
    print("# Static SASS mix of the product kernels (cuobjdump -sass, per thread, whole kernel)\n")
    print("`UTMALDG`/`UTMASTG` = `cp.async.bulk.tensor` (TMA), `SYNCS` = mbarrier, `FADD2`/`FFMA2` = packed f32x2, `SHFL` = warp shuffle.\n")
    hdr = "| instance | total | " + " | ".join(KEYS) + " |\n|---|---|" + "---|" * len(KEYS)
    lib = os.path.join(ROOT, "smfft_b200", "lib", "libsmfft.so")
    rows = []
    tot = collections.Counter()
    for name, c in functions(lib):
        for k in KEYS:
            tot[k] += c[k]
        m = re.search(r"BlockCfgILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E.*?Li(\d+)EEELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(-?\d+|n\d+)E", name)
        if not m:
            continue
        e, b, f, d, r, tw, arith, mode, io, st, reps, minb, pf = m.groups()
        if int(d) != (1 if int(mode) == 2 else 0) or int(tw) != 0:
            continue
        if int(reps) == 3:
            continue
        label = f"{MODES[int(mode)]} N={1 << int(e)} {'reorder' if int(r) else 'no-reorder'} R={1 << int(b)} F={f} io={IOS[int(io)]} reps={reps} arith={arith} minb={minb}"
        rows.append(((int(mode), int(reps), int(e), -int(r), int(io)), row(label, c)))
    big = []
    for name, c in functions(lib):
        m = re.search(r"big_pass_kernelILi(\d+)ELi(\d+)ELi(\d+)E", name)
        if m and int(m.group(2)) == 0:
            big.append((int(m.group(1)), int(m.group(3)), row(f"two-pass C2C, pass {'AB'[int(m.group(3))]}, {1 << int(m.group(1))}-point block transforms x 16 per tile", c)))
    print("## libsmfft.so -- totals over ALL instances: " + ", ".join(f"{k} {tot[k]}" for k in ("UTMALDG", "UTMASTG", "SYNCS", "FADD2", "SHFL")) + "\n")
    print("Forward (C2R: inverse), table twiddles; every staging variant compiled into the library:\n")
    print(hdr)
    for _, r_ in sorted(rows):
        print(r_)
    if big:
        print("\n## two-pass transforms of 2^15 .. 2^18 points (csrc/big_fft.cu, forward): TMA box in, TMA box out\n")
        print(hdr)
        for _, _, r_ in sorted(big):
            print(r_)
    so = os.path.join(ROOT, "tests", "compat", "_build", "libcompat_kernels.so")
    if os.path.exists(so):
        print("\n## reference-contract wrapper kernels (include/smfft/compat.cuh) and the native device primitive (include/smfft/device.cuh)\n")
        print(hdr)
        out = []
        for name, c in functions(so):
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            if re.search(r"(SMFFT_DIT_(external|multiple)<FFT_\d+_forward(_noreorder)?>|FFT_GPU_external<|FFT_GPU_R2C_C2R_external<.*FFT_forward|native_convolve<|user_convolve<)", dem):
                out.append((dem, row("`" + dem.split("(")[0].replace("void ", "") + "`", c)))
        for _, r_ in sorted(out):
            print(r_)


main()
