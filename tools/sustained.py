#!/usr/bin/env python
"""Sustained (power-capped) timing per size: `count` back-to-back launches of one configuration between
two events, SM clock sampled with nvidia-smi meanwhile (measurement tool)."""
import json
import os
import statistics
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
COUNT = int(sys.argv[2]) if len(sys.argv) > 2 else 150
x = torch.rand((PTS, 2), device="cuda")
y = torch.empty_like(x)
xr = x.view(-1)
out = {}


def clocks():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
    try:
        a, b = r.stdout.strip().split(",")
        return float(a), float(b)
    except Exception:
        return None, None


def sustained(fn):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(COUNT // 2):
        fn()
    mhz, watts = clocks()   # sampled mid-run (queue is deep enough to keep the GPU busy)
    for _ in range(COUNT - COUNT // 2):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / COUNT, 4), mhz, watts


for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    for reorder in (1, 0):
        for tw in (0, 1):
            sm.set_option("twiddle", tw)
            ms, mhz, w = sustained(lambda: sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder)))
            out[f"c2c_{n}_{'r' if reorder else 'n'}_{'lut' if tw == 0 else 'mufu'}"] = {"ms": ms, "sm_mhz": mhz, "watts": w}
            print(n, reorder, tw, ms, mhz, w, flush=True)
        time.sleep(0.3)
sm.set_option("twiddle", 0)
for n in (64, 256, 512, 1024, 2048, 4096, 8192):
    for inv in (0, 1):
        ms, mhz, w = sustained(lambda: sm.exec_r2c_c2r(xr, y, n, 2 * PTS // n, inv))
        out[f"{'c2r' if inv else 'r2c'}_{n}"] = {"ms": ms, "sm_mhz": mhz, "watts": w}
        print("real", n, inv, ms, mhz, w, flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sustained.json", "w"), indent=1)
