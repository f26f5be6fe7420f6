#!/usr/bin/env python
"""cuFFT launches for an ncu look at how the library streams a batch (reported baseline only)."""
import ctypes
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pts = 1 << 27
x = torch.rand((pts, 2), device="cuda")
y = torch.empty_like(x)
lib = ctypes.CDLL("libcufft.so.11")
h = ctypes.c_int(0)
assert lib.cufftPlan1d(ctypes.byref(h), n, 0x29, pts // n) == 0
for _ in range(3):
    lib.cufftExecC2C(h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), -1)
torch.cuda.synchronize()
