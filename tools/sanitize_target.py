#!/usr/bin/env python
"""Small all-modes launch sequence for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm
from oracle import oracle_np as O

bad = 0
for io in (0, 1):
    sm.set_option("io", io)
    for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
        nf = 3 * (8192 // n) + 1
        x = O.uniform_c64(nf, n)
        dx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2)).cuda()
        dy = torch.zeros_like(dx)
        for inverse, reorder in ((0, 1), (1, 0)):
            sm.exec_c2c(dx, dy, n, nf, bool(inverse), bool(reorder))
            torch.cuda.synchronize()
            err = O.rel_l2(dy.cpu().numpy().view(np.complex64).reshape(nf, n), O.ct_c2c_fp64(x, bool(inverse), bool(reorder)))
            bad += err > 1e-5
        xr = O.uniform_f32(nf, 2 * n)
        dr = torch.from_numpy(xr).cuda()
        dc = torch.zeros((nf, n, 2), dtype=torch.float32, device="cuda")
        sm.exec_r2c_c2r(dr, dc, 2 * n, nf, 0)
        sm.exec_r2c_c2r(dc, dr, 2 * n, nf, 1)
        torch.cuda.synchronize()
        bad += O.rel_l2(dr.cpu().numpy() / n, xr) > 1e-5
x = torch.rand((100 * 4096, 2), device="cuda")
y = torch.empty_like(x)
sm.FFT_multiple_benchmark(x, y, 1024, 400, False, True)
sm.FFT_multiple_benchmark(x, y, 32, 12800, False, False)
print("sanitize_target done, mismatches:", int(bad))
