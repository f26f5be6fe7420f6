#!/usr/bin/env python
"""tools/sustained_ctas.py -- the 16-launch bench step (sustained load) with the CTAs-per-SM of one size overridden (measurement tool)."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
sm.FFT_init()
SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096]


def run(target, ctas, steps=12):
    per = {}
    for s in range(steps + 3):
        for n in SIZES:
            for reorder in (1, 0):
                sm.set_option("ctas_per_sm", ctas if n == target else 0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder))
                e1.record()
                if s >= 3:
                    per.setdefault((n, reorder), []).append((e0, e1))
    torch.cuda.synchronize()
    sm.set_option("ctas_per_sm", 0)
    med = {k: statistics.median(a.elapsed_time(b) for a, b in v) for k, v in per.items()}
    return round(med[(target, 1)], 4), round(med[(target, 0)], 4), round(sum(med.values()) / len(med), 4)


for target, opts in ((4096, (0, 2, 3, 0)), (2048, (0, 4, 6, 0)), (1024, (0, 1, 3, 0)), (512, (0, 1, 3, 0))):
    print(target, {c: run(target, c) for c in opts}, flush=True)
