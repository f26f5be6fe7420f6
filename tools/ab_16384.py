#!/usr/bin/env python
"""16384-point C2C: TMA store path (io = 2) against TMA in / registers out with the single buffer refilled behind the final
exchange (io = 3); interleaved, 4 GiB batches, both orders.  Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

N = 16384
nf = (1 << 29) // N
x = torch.rand((nf, N, 2), device="cuda")
y = torch.empty_like(x)
out = {"n": N, "n_ffts": nf, "unit": "ms per 4 GiB batch (in + out 8 GiB)"}
for reorder in (True, False):
    res = {2: [], 3: []}
    for rnd in range(5):
        for io in (2, 3):
            sm.set_option("io", io)
            for _ in range(3):
                sm.exec_c2c(x, y, N, nf, False, reorder)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                sm.exec_c2c(x, y, N, nf, False, reorder)
            ev[1].record()
            torch.cuda.synchronize()
            res[io].append(ev[0].elapsed_time(ev[1]) / 10)
    out["natural" if reorder else "noreorder"] = {"tma_store": round(min(res[2]), 4), "regs_out_single_buffer": round(min(res[3]), 4),
                                                 "all_tma": [round(v, 4) for v in res[2]], "all_regs": [round(v, 4) for v in res[3]]}
sm.set_option("io", 0)
print(json.dumps(out))
