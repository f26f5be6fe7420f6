#!/usr/bin/env python
"""tools/carveout_step_probe.py -- does the register-direct path lose inside the bench step because consecutive kernels use
DIFFERENT shared-memory carve-outs (the SM partition is reconfigured between them), rather than because of the power cap?
Runs the sustained 16-launch step with the static table and with the register-direct shapes (io = 4) under several
process-wide carve-outs ("carveout" option: -2 = per kernel [TMA kernels max shared, register-direct driver default],
else percent of 228 KB for EVERY kernel), plus a same-kernel burst (30 back-to-back launches of one instance).

    python tools/carveout_step_probe.py [out.json] [steps]
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smfft_b200 as sm  # noqa: E402

PTS = 1 << 29
SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096]


def step(x, y, rec):
    for n in SIZES:
        for reorder in (1, 0):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder))
            e1.record()
            rec.setdefault(f"{n}{'r' if reorder else 'n'}", []).append((e0, e1))


def burst(x, y, n, count=30):
    for _ in range(3):
        sm.exec_c2c(x, y, n, PTS // n, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(count):
        sm.exec_c2c(x, y, n, PTS // n, False, True)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / count, 4)


def main(out_path, steps):
    torch.cuda.set_device(0)
    sm.FFT_init()
    x = torch.rand((PTS, 2), device="cuda")
    y = torch.empty_like(x)
    out = {}
    for carve in (-2, 72, 58, 44, 100):
        sm.set_option("carveout", carve)
        row = {}
        for name, io in (("static", 0), ("reg_a", 4)):
            sm.set_option("io", io)
            step(x, y, {})
            step(x, y, {})
            torch.cuda.synchronize()
            rec = {}
            for _ in range(steps):
                step(x, y, rec)
            torch.cuda.synchronize()
            med = {k: round(statistics.median(a.elapsed_time(b) for a, b in v), 4) for k, v in rec.items()}
            row[name] = {"per_size_ms": med, "step_ms": round(sum(med.values()), 4),
                         "burst30_ms": {str(n): burst(x, y, n) for n in (128, 256, 512, 1024)}}
        out[str(carve)] = row
        print(carve, {k: (v["step_ms"], v["burst30_ms"], {s: v["per_size_ms"][s] for s in ("128r", "256r", "512r", "1024r")}) for k, v in row.items()}, flush=True)
    sm.set_option("carveout", -2)
    sm.set_option("io", 0)
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/carveout_step_probe.json", int(sys.argv[2]) if len(sys.argv) > 2 else 8)
