#!/usr/bin/env python
"""tools/ab.py LIB_A LIB_B [out.json] -- interleaved A/B timing of two builds of libsmfft on one B200.

Both libraries are loaded in one process and timed alternately (A, B, A, B ...) on the same 4 GiB batch so that
clock / thermal drift hits both equally (box-to-box and minute-to-minute variance is larger than most tuning
deltas).  Measurement tool only.
"""
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
    lib.smfft_exec_c2c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int,
                                   ctypes.c_int]
    lib.smfft_multiple_benchmark.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int,
                                             ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.smfft_r2c_multiple_benchmark.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
                                                 ctypes.POINTER(ctypes.c_double)]
    lib.smfft_exec_r2c_c2r.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int]
    assert lib.smfft_init() == 0
    return lib


def once(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert fn() == 0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def ab(fa, fb, reps=15):
    for _ in range(3):
        fa(), fb()
    ta, tb = [], []
    for _ in range(reps):
        ta.append(once(fa))
        tb.append(once(fb))
    return round(statistics.median(ta), 4), round(statistics.median(tb), 4)


def main():
    A, B = load(sys.argv[1]), load(sys.argv[2])
    xi, yo = x.data_ptr(), y.data_ptr()
    out = {}
    sizes = [int(s) for s in sys.argv[4].split(",")] if len(sys.argv) > 4 else [256, 512, 1024, 2048, 4096]
    for n in sizes:
        row = {}
        for reorder in (1, 0):
            row[f"c2c_r{reorder}"] = ab(lambda: A.smfft_exec_c2c(xi, yo, n, PTS // n, 0, reorder),
                                        lambda: B.smfft_exec_c2c(xi, yo, n, PTS // n, 0, reorder))
        for inv in (0, 1):
            row["c2r" if inv else "r2c"] = ab(lambda: A.smfft_exec_r2c_c2r(xi, yo, 2 * n, PTS // n, inv),
                                              lambda: B.smfft_exec_r2c_c2r(xi, yo, 2 * n, PTS // n, inv))
        ms = ctypes.c_double(0)
        row["multiple_r1"] = ab(lambda: A.smfft_multiple_benchmark(xi, yo, n, PTS // n, 0, 1, ctypes.byref(ms)),
                                lambda: B.smfft_multiple_benchmark(xi, yo, n, PTS // n, 0, 1, ctypes.byref(ms)), reps=7)
        row["r2c_multiple"] = ab(lambda: A.smfft_r2c_multiple_benchmark(xi, yo, 2 * n, PTS // n, ctypes.byref(ms)),
                                 lambda: B.smfft_r2c_multiple_benchmark(xi, yo, 2 * n, PTS // n, ctypes.byref(ms)), reps=7)
        out[n] = row
        print(n, row, flush=True)
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], "w"), indent=1)


main()
