#!/usr/bin/env python
"""tools/carveout_probe.py -- product kernels timed with different shared-memory carve-outs (measurement tool).
-2 = product setting (max shared), -1 = driver default, else percent.  The LDG/STG paths like the L1 a small carve-out leaves
(tools/fftlike_copy.cu); do the TMA-in / register-out kernels?"""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
sm.FFT_init()


def t(fn):
    fn(); fn()
    return round(statistics.median([fn() for _ in range(9)]), 4)


for cv in (-2, -1, 75, 50, -2):
    sm.set_option("carveout", cv)
    row = {}
    for n in (32, 128, 256, 1024, 4096):
        row[f"c2c{n}"] = t(lambda: sm.FFT_external_benchmark(x, y, n, PTS // n, False, True))
    for n in (1024, 2048, 4096):
        row[f"r2c{n}"] = t(lambda: sm.R2C_C2R_external_benchmark(x, y, n, 2 * PTS // n, 0))
        row[f"c2r{n}"] = t(lambda: sm.R2C_C2R_external_benchmark(x, y, n, 2 * PTS // n, 1))
    for n in (256, 2048):
        row[f"mult{n}"] = t(lambda: sm.FFT_multiple_benchmark(x, y, n, PTS // n, False, True))
    print("carveout", cv, row, flush=True)
