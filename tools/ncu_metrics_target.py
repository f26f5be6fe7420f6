#!/usr/bin/env python
"""Launch list for the per-size DRAM-traffic / pipe-utilisation capture (tools/gpu/r2_h.sh): one FFT_external launch per
bench configuration on the bench's own 4 GiB batch, then one FFT_multiple launch per configuration, then R2C / C2R."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    for reorder in (1, 0):
        sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder))
for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    for reorder in (1, 0):
        sm.FFT_multiple_benchmark(x, y, n, PTS // n, False, bool(reorder))
xr = x.view(-1)
for n in (512, 1024, 2048, 4096):
    sm.exec_r2c_c2r(xr, y, n, 2 * PTS // n, 0)
    sm.exec_r2c_c2r(xr, y, n, 2 * PTS // n, 1)
torch.cuda.synchronize()
