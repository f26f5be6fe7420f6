#!/usr/bin/env python
"""Launch sequence for ncu captures of the reference-contract wrapper kernels: this library's (include/smfft/compat.cuh,
built from tests/compat/compat_kernels.cu) or the reference's own (oracle/_ref), on a 1 GiB batch.
usage: compat_target.py <compat|reference> <external|multiple> <N> <inverse> <reorder>"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refkernels as R  # noqa: E402
from tools.compat_bench import load_compat  # noqa: E402

who, kind, n, inverse, reorder = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
pts = 1 << 27
x = torch.rand((pts, 2), device="cuda")
y = torch.empty_like(x)
lib = load_compat() if who == "compat" else None
for _ in range(4):
    if who == "compat":
        (lib.compat_ct_external if kind == "external" else lib.compat_ct_multiple)(x.data_ptr(), y.data_ptr(), n, pts // n, inverse, reorder)
    else:
        (R.ct_external if kind == "external" else R.ct_multiple)(x, y, n, pts // n, inverse, reorder)
torch.cuda.synchronize()
