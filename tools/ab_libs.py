#!/usr/bin/env python
"""tools/ab_libs.py out.json N[,N..] LIB [LIB ...] -- interleaved timing of several builds of libsmfft on the same 4 GiB batch:
smfft_exec_c2c for the given sizes, both orders, CUDA events per launch, rounds A B C A B C ..., median.  With
`--burst K` every measurement is K back-to-back launches (steady-state behaviour), else single launches."""
import ctypes
import json
import statistics
import sys

import torch

args = sys.argv[1:]
burst = 1
if "--burst" in args:
    i = args.index("--burst")
    burst = int(args[i + 1])
    del args[i:i + 2]
io = 0
if "--io" in args:
    i = args.index("--io")
    io = int(args[i + 1])
    del args[i:i + 2]
out_path, sizes, libs = args[0], [int(s) for s in args[1].split(",")], args[2:]
PTS = 1 << 29
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)


def load(path):
    lib = ctypes.CDLL(path)
    lib.smfft_exec_c2c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    lib.smfft_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    assert lib.smfft_init() == 0
    lib.smfft_set_option(b"io", io)
    return lib


L = [load(p) for p in libs]
out = {"libs": libs, "burst": burst, "io": io, "ms": {}}
for n in sizes:
    for reorder in (1, 0):
        ts = [[] for _ in L]
        for r in range(13):
            for j, lib in enumerate(L):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(burst):
                    assert lib.smfft_exec_c2c(x.data_ptr(), y.data_ptr(), n, PTS // n, 0, reorder) == 0
                e1.record()
                torch.cuda.synchronize()
                if r >= 3:
                    ts[j].append(e0.elapsed_time(e1) / burst)
        key = f"{n}{'r' if reorder else 'n'}"
        out["ms"][key] = [round(statistics.median(t), 4) for t in ts]
        print(key, out["ms"][key], flush=True)
json.dump(out, open(out_path, "w"), indent=1)
