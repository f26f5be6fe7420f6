#!/usr/bin/env python
"""cuFFT launches for an ncu look at how the library streams a batch (reported baseline only).
usage: cufft_target.py [N ...]   (default: every size 32..4096; 1 GiB batch, three launches per size)"""
import ctypes
import sys

import torch

sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 256, 512, 1024, 2048, 4096]
pts = 1 << 27
x = torch.rand((pts, 2), device="cuda")
y = torch.empty_like(x)
lib = ctypes.CDLL("libcufft.so.11")
for n in sizes:
    h = ctypes.c_int(0)
    assert lib.cufftPlan1d(ctypes.byref(h), n, 0x29, pts // n) == 0
    for _ in range(3):
        lib.cufftExecC2C(h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), -1)
    torch.cuda.synchronize()
    lib.cufftDestroy(h)
