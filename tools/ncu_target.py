#!/usr/bin/env python
"""Small launch sequence for ncu captures on a 1 GiB batch (>> L2).
usage: ncu_target.py <kind> <N> [reorder]   kind = c2c | r2c | c2r | multiple"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

kind = sys.argv[1] if len(sys.argv) > 1 else "c2c"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reorder = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pts = 1 << int(os.environ.get("SMFFT_NCU_LOG2_POINTS", "27"))
sm.set_option("io", int(os.environ.get("SMFFT_IO", "0")))   # 4 / 5: the register-direct alternates
x = torch.rand((pts, 2), device="cuda")
y = torch.empty_like(x)
for _ in range(4):
    if kind == "c2c":
        sm.exec_c2c(x, y, n, pts // n, False, bool(reorder))
    elif kind == "r2c":
        sm.exec_r2c_c2r(x, y, n, 2 * pts // n, 0)
    elif kind == "c2r":
        sm.exec_r2c_c2r(x, y, n, 2 * pts // n, 1)
    else:
        sm.FFT_multiple_benchmark(x, y, n, pts // n, False, bool(reorder))
torch.cuda.synchronize()
