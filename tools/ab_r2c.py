#!/usr/bin/env python
"""tools/ab_r2c.py LIB... -- interleaved timing of R2C (and C2R) for several builds x staging options (measurement tool)."""
import ctypes, statistics, sys
import torch

PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)


def load(path):
    lib = ctypes.CDLL(path)
    lib.smfft_exec_r2c_c2r.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int]
    lib.smfft_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    assert lib.smfft_init() == 0
    return lib


libs = [(p.split("libsmfft")[-1].replace(".so", "") or "product", load(p)) for p in sys.argv[1:]]
xi, yo = x.data_ptr(), y.data_ptr()
for n in (2048, 8192, 4096):
    for inv in (0, 1):
        cases = [(name, lib, io) for name, lib in libs for io in (0, 3)]
        ts = {(name, io): [] for name, _, io in cases}
        for rep in range(14):
            for name, lib, io in cases:
                lib.smfft_set_option(b"io", io)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                assert lib.smfft_exec_r2c_c2r(xi, yo, n, 2 * PTS // n, inv) == 0
                e1.record()
                torch.cuda.synchronize()
                if rep >= 2:
                    ts[(name, io)].append(e0.elapsed_time(e1))
        print(n, "c2r" if inv else "r2c", {f"{k[0]}_io{k[1]}": round(statistics.median(v), 4) for k, v in ts.items()}, flush=True)
