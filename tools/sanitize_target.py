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

# round 2: alternates (register-direct shapes A / B, alternate TMA shapes) incl. the first-use selection, the repeated path
# with three repetitions, the reference-contract wrapper kernels (warp-shuffle engines) and the native device primitive
import ctypes

from tests.compat.build_compat import build as build_compat


def run(n, nf, inverse, reorder):
    global bad
    x = O.uniform_c64(nf, n)
    dx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2)).cuda()
    dy = torch.zeros_like(dx)
    sm.exec_c2c(dx, dy, n, nf, inverse, reorder)
    torch.cuda.synchronize()
    bad += O.rel_l2(dy.cpu().numpy().view(np.complex64).reshape(nf, n), O.ct_c2c_fp64(x, inverse, reorder)) > 1e-5


for io in (4, 5):
    sm.set_option("io", io)
    for n in (128, 256, 512, 1024):
        run(n, 3 * (8192 // n) + 5, False, True)
sm.set_option("io", 0)
sm.set_option("select", 1)
sm.set_option("select_min_log2_points", 12)
for n in (32, 128, 1024, 4096):
    run(n, 3 * (8192 // n) + 5, False, True)
sm.set_option("select", 0)
sm.set_option("select_min_log2_points", 24)
sm.set_option("select_reset", 1)
for n in (32, 1024, 4096):
    nf = 2 * (8192 // n) + 3
    x = torch.rand((nf, n, 2), device="cuda")
    y = torch.empty_like(x)
    sm.exec_repeated(x, y, n, nf, False, True, 0, 3)
    sm.exec_repeated(x, y, n, nf, False, False, 0, 3)
# 16384 reals (8192-point core, R2C and C2R) with more tiles than SMs
xr = O.uniform_f32(2 * 148 + 3, 16384)
dr = torch.from_numpy(xr).cuda()
dc = torch.zeros((xr.shape[0], 8192, 2), dtype=torch.float32, device="cuda")
sm.exec_r2c_c2r(dr, dc, 16384, xr.shape[0], 0)
sm.exec_r2c_c2r(dc, dr, 16384, xr.shape[0], 1)
torch.cuda.synchronize()
bad += O.rel_l2(dr.cpu().numpy() / 8192, xr) > 1e-5
# 8192 / 16384 points with more tiles than SMs: every persistent CTA refills its buffers (16384: the ONE buffer, behind the
# final exchange, while the results leave from registers)
for n, nf in ((8192, 3 * 148 + 7), (16384, 2 * 148 + 9)):
    for io in (0, 2, 3):
        sm.set_option("io", io)
        run(n, nf, False, True)
        run(n, nf, True, False)
sm.set_option("io", 0)
# multi-pass transforms (2^15 .. 2^24 points): strided TMA boxes in and out, chunked scratch
sm.set_option("two_pass_chunk_mib", 2)
for e in (15, 16, 17, 18, 19, 20, 21, 22):   # 19 / 20: 1024-point passes on 8-transform tiles; 21 and up: three passes, the last one gathering its 16 transforms box by box
    run(1 << e, 3 if e <= 18 else 2, False, True)
    run(1 << e, 2 if e <= 18 else 1, True, True)
sm.set_option("two_pass_chunk_mib", 1024)
lib = ctypes.CDLL(build_compat())
P, I = ctypes.c_void_p, ctypes.c_int
lib.compat_ct_external.argtypes = [P, P, I, I, I, I]
lib.compat_ct_multiple.argtypes = [P, P, I, I, I, I]
lib.native_fft_launch.argtypes = [P, P, I, I, I, I, P]
lib.native_convolve_launch.argtypes = [P, P, P, I, I, I, P]
tw = sm.twiddle_table()
for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    x = torch.rand((400, n, 2), device="cuda")
    y = torch.empty_like(x)
    for inverse in (0, 1):
        for reorder in (1, 0):
            assert lib.compat_ct_external(x.data_ptr(), y.data_ptr(), n, 400, inverse, reorder) == 0
    assert lib.compat_ct_multiple(x.data_ptr(), y.data_ptr(), n, 400, 0, 0) == 0
    assert lib.compat_ct_multiple(x.data_ptr(), y.data_ptr(), n, 400, 0, 1) == 0
    assert lib.native_fft_launch(x.data_ptr(), y.data_ptr(), n, 384, 0, 1, tw) == 0
    if n in (256, 1024, 4096):
        h = torch.rand((n, 2), device="cuda")
        assert lib.native_convolve_launch(x.data_ptr(), h.data_ptr(), y.data_ptr(), n, 384, 0, tw) == 0
torch.cuda.synchronize()
print("sanitize_target done, mismatches:", int(bad))
