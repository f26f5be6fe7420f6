#!/usr/bin/env python
"""Sweep CTAs/SM (the persistent grid size) of the product kernels on a 4 GiB batch (measurement tool)."""
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
out = {}
for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    for reorder in (1, 0):
        row = {}
        for per in (0, 1, 2, 3, 4, 5, 6, 8):
            sm.set_option("ctas_per_sm", per)
            try:
                for _ in range(3):
                    sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder))
                torch.cuda.synchronize()
                ts = []
                for _ in range(9):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder))
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                row[per] = round(statistics.median(ts), 4)
            except Exception as ex:
                row[per] = str(ex)[:60]
                break
        out[f"{n}{'r' if reorder else 'n'}"] = row
        print(n, reorder, row, flush=True)
sm.set_option("ctas_per_sm", 0)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep_ctas.json", "w"), indent=1)
