#!/usr/bin/env python
"""tools/staging_probe.py -- TMA loads + TMA stores (io = 2) against TMA loads + stores from registers (io = 3), per transform,
each kernel in its own steady state (COUNT back-to-back launches, mean of the last 20) and in short bursts (first 5): the
table behind Tuning::STG / STG_R2C / STG_C2R, re-derived under the round-2 protocol.

    python tools/staging_probe.py [out.json] [count]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smfft_b200 as sm  # noqa: E402

PTS = 1 << 29


def timeline(fn, count):
    fn()
    torch.cuda.synchronize()
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(count + 1)]
    ev[0].record()
    for i in range(count):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(count)]
    return round(sum(ts[:5]) / 5, 4), round(sum(ts[-20:]) / 20, 4)


def main(out_path, count):
    torch.cuda.set_device(0)
    sm.FFT_init()
    x = torch.rand((PTS, 2), device="cuda")
    y = torch.empty_like(x)
    xr = x.view(-1)
    out = {}
    cases = [(f"c2c_{n}{'r' if r else 'n'}", (lambda n=n, r=r: sm.exec_c2c(x, y, n, PTS // n, False, bool(r)))) for n in (32, 64, 128, 256, 512, 1024, 2048, 4096) for r in (1, 0)]
    cases += [(f"{'c2r' if inv else 'r2c'}_{n}", (lambda n=n, inv=inv: sm.exec_r2c_c2r(xr, y, n, 2 * PTS // n, inv))) for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192) for inv in (0, 1)]
    for name, fn in cases:
        row = {}
        for io in (0, 2, 3):
            sm.set_option("io", io)
            row[{0: "default", 2: "tma", 3: "tma_stg"}[io]] = timeline(fn, count)
        out[name] = row
        print(name, row, flush=True)
    sm.set_option("io", 0)
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/staging_probe.json", int(sys.argv[2]) if len(sys.argv) > 2 else 60)
