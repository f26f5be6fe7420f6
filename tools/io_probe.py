#!/usr/bin/env python
"""tools/io_probe.py -- R2C / C2R external timings per staging option (io = 0 auto, 2 TMA, 3 TMA in + register out)."""
import statistics
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
sm.FFT_init()
for n in (2048, 4096, 8192):
    row = {}
    for io in (0, 2, 3):
        sm.set_option("io", io)
        for inv, name in ((0, "r2c"), (1, "c2r")):
            ts = [sm.R2C_C2R_external_benchmark(x, y, n, 2 * PTS // n, inv) for _ in range(12)][2:]
            row[f"{name}_io{io}"] = round(statistics.median(ts), 4)
    print(n, row, flush=True)
