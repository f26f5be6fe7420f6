#!/usr/bin/env python
"""tools/real_ctas_probe.py -- R2C / C2R with the CTAs-per-SM overridden, interleaved (measurement tool)."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
sm.FFT_init()
for n, opts in ((2048, (0, 5, 3)), (1024, (0, 5, 3)), (4096, (0, 5, 4)), (8192, (0, 2, 4))):
    for inv, name in ((0, "r2c"), (1, "c2r")):
        ts = {c: [] for c in opts}
        for rep in range(14):
            for c in opts:
                sm.set_option("ctas_per_sm", c)
                t = sm.R2C_C2R_external_benchmark(x, y, n, 2 * PTS // n, inv)
                if rep >= 2:
                    ts[c].append(t)
        print(n, name, {c: round(statistics.median(v), 4) for c, v in ts.items()}, flush=True)
sm.set_option("ctas_per_sm", 0)
