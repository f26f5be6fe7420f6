#!/usr/bin/env python
"""tools/ab_multiple.py LIB_A LIB_B out.json -- interleaved A/B of FFT_multiple (C2C, both reorder modes) for the small sizes."""
import ctypes
import json
import statistics
import sys

import torch

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)


def load(path):
    lib = ctypes.CDLL(path)
    lib.smfft_multiple_benchmark.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_double)]
    assert lib.smfft_init() == 0
    return lib


A, B = load(sys.argv[1]), load(sys.argv[2])
out = {"A": sys.argv[1], "B": sys.argv[2], "protocol": "interleaved A B A B, library-timed CUDA events, median of 15", "ms": {}}
for n in [int(a) for a in sys.argv[4].split(",")] if len(sys.argv) > 4 else (32, 64, 128):
    for reorder in (1, 0):
        ta, tb = [], []
        for i in range(18):
            for lib, ts in ((A, ta), (B, tb)):
                ms = ctypes.c_double(0)
                assert lib.smfft_multiple_benchmark(x.data_ptr(), y.data_ptr(), n, PTS // n, 0, reorder, ctypes.byref(ms)) == 0
                if i >= 3:
                    ts.append(ms.value)
        a, b = statistics.median(ta), statistics.median(tb)
        out["ms"][f"{n}{'r' if reorder else 'n'}"] = {"A": round(a, 4), "B": round(b, 4), "B_over_A": round(b / a, 4)}
        print(n, reorder, out["ms"][f"{n}{'r' if reorder else 'n'}"], flush=True)
json.dump(out, open(sys.argv[3], "w"), indent=1)
