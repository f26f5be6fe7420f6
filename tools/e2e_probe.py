#!/usr/bin/env python
"""tools/e2e_probe.py -- host-to-host pipeline (smfft_pipeline_host) throughput vs chunk size (measurement tool)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smfft_b200 as sm

PTS = 1 << 29
sm.FFT_init()
hx = torch.empty((PTS, 2), dtype=torch.float32, pin_memory=True)
hy = torch.empty((PTS, 2), dtype=torch.float32, pin_memory=True)
hx.uniform_()
n = 1024
for mib in (128, 64, 32, 16, 8, 32, 128):
    chunk = (mib << 20) // (n * 8)
    sm.pipeline_host(hx, hy, n, PTS // n, False, True, 0, chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        sm.pipeline_host(hx, hy, n, PTS // n, False, True, 0, chunk)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk {mib:4d} MiB: {dt * 1e3:8.2f} ms per 4 GiB batch, {PTS * 16 / dt / 1e9:6.1f} GB/s (in + out)", flush=True)
