#!/usr/bin/env python
"""Small launch sequence for ncu captures: a few FFT_external launches on a 1 GiB batch (>> L2)."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import smfft_b200 as sm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reorder = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pts = 1 << 27
x = torch.rand((pts, 2), device="cuda")
y = torch.empty_like(x)
for _ in range(4):
    sm.exec_c2c(x, y, n, pts // n, False, bool(reorder))
torch.cuda.synchronize()
