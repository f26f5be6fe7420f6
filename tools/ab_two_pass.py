#!/usr/bin/env python
"""tools/ab_two_pass.py [chunk_mib] -- two-pass transforms (2^15 .. 2^18 points) on a 4 GiB batch: ms per launch (10 back-to-back,
best of 3), values checked on the first rows against torch.fft (tool only)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sm.set_option("two_pass_chunk_mib", chunk)
out = {"two_pass_chunk_mib": chunk}
for e in ([int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else (15, 16, 17, 18)):
    n = 1 << e
    nf = (1 << 29) // n
    x = torch.rand((nf, n, 2), device="cuda")
    y = torch.empty_like(x)
    best = 1e9
    for _ in range(3):
        for _ in range(2):
            sm.exec_c2c(x, y, n, nf, False, True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(10):
            sm.exec_c2c(x, y, n, nf, False, True)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]) / 10)
    ref = torch.fft.fft(torch.view_as_complex(x[:2]))
    err = (torch.view_as_complex(y[:2]) - ref).norm() / ref.norm()
    out[str(n)] = {"ms": round(best, 4), "frac_of_floor": round((2 if e <= 20 else 3) * 1.313 / best, 3), "rel_l2_vs_torch": float(err)}
    del x, y
print(json.dumps(out))
