#!/usr/bin/env python
"""tools/compat_prefetch_probe.py -- SMFFT_DIT_external<P> (compat.cuh) with different L2 prefetch distances vs the reference's
kernel: interleaved, CUDA events per launch, median.  Variant libraries: tests/compat/_build/libcompat_kernels_pf{0,2,32}.so
(SMFFT_COMPAT_PREFETCH_BYTES = 0 / 2 / 32 MiB) next to the default build (8 MiB)."""
import ctypes
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refkernels as R  # noqa: E402

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
B = os.path.join(ROOT, "tests", "compat", "_build")
libs = {"pf0": "libcompat_kernels_pf0.so", "pf2": "libcompat_kernels_pf2.so", "pf8": "libcompat_kernels.so", "pf32": "libcompat_kernels_pf32.so"}
L = {}
for k, f in libs.items():
    if os.path.exists(os.path.join(B, f)):
        L[k] = ctypes.CDLL(os.path.join(B, f))
        L[k].compat_ct_external.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
out = {}
for n in (32, 64, 128, 256, 1024, 4096):
    for reorder in (1, 0):
        arms = {k: (lambda lib=lib: lib.compat_ct_external(x.data_ptr(), y.data_ptr(), n, PTS // n, 0, reorder)) for k, lib in L.items()}
        arms["reference"] = lambda: R.ct_external(x, y, n, PTS // n, False, reorder)
        ts = {k: [] for k in arms}
        for r in range(10):
            for k, fn in arms.items():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                if r >= 3:
                    ts[k].append(e0.elapsed_time(e1))
        key = f"{n}{'r' if reorder else 'n'}"
        out[key] = {k: round(statistics.median(v), 4) for k, v in ts.items()}
        print(key, out[key], flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/compat_prefetch_probe.json", "w"), indent=1)
