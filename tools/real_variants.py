#!/usr/bin/env python
"""R2C / C2R: TMA-store (io=2) vs register-store (io=3) staging and CTAs/SM, 4 GiB real batch (measurement tool)."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
xr = x.view(-1)
out = {}


def t(fn, reps=9):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return round(statistics.median(ts), 4)


for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192):
    row = {}
    for io in (2, 3):
        for per in (2, 3):
            sm.set_option("io", io)
            sm.set_option("ctas_per_sm", per)
            row[f"r2c_io{io}_cta{per}"] = t(lambda: sm.exec_r2c_c2r(xr, y, n, 2 * PTS // n, 0))
            row[f"c2r_io{io}_cta{per}"] = t(lambda: sm.exec_r2c_c2r(xr, y, n, 2 * PTS // n, 1))
    out[n] = row
    print(n, row, flush=True)
sm.set_option("io", 0)
sm.set_option("ctas_per_sm", 0)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/real_variants.json", "w"), indent=1)
