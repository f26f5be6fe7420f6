#!/usr/bin/env python
"""tools/burst_timeline.py -- per-launch time of COUNT back-to-back launches of one kernel (events between all of them):
where along a sustained burst does a shape slow down?  Arms: the static table's instance, register-direct A (io = 4), cuFFT.
Prints one line per (arm, N) with the per-launch milliseconds; SM clock / power sampled at the end of each burst.

    [SMFFT_LIB=path] python tools/burst_timeline.py [out.json] [count]
"""
import ctypes
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smfft_b200 as sm  # noqa: E402

PTS = 1 << 29


def clocks():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
    return r.stdout.strip()


def timeline(fn, count):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(count + 1)]
    ev[0].record()
    for i in range(count):
        fn()
        ev[i + 1].record()
    c = clocks()   # sampled while the queue is still draining
    torch.cuda.synchronize()
    return [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(count)], c


def main(out_path, count):
    torch.cuda.set_device(0)
    sm.FFT_init()
    x = torch.rand((PTS, 2), device="cuda")
    y = torch.empty_like(x)
    cu = ctypes.CDLL("libcufft.so.11")
    out = {"lib": os.environ.get("SMFFT_LIB", "product"), "count": count}
    for n in (128, 256, 512, 1024, 2048, 4096):
        h = ctypes.c_int(0)
        assert cu.cufftPlan1d(ctypes.byref(h), n, 0x29, PTS // n) == 0
        arms = {"tma": (2 if n == 256 else 0, lambda: sm.exec_c2c(x, y, n, PTS // n, False, True)),   # 256 points: the default is register-direct, io = 2 is its TMA instance
                "reg_a": (4, lambda: sm.exec_c2c(x, y, n, PTS // n, False, True)),
                "reg_b": (5, lambda: sm.exec_c2c(x, y, n, PTS // n, False, True)),
                "cufft": (0, lambda: cu.cufftExecC2C(h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), -1))}
        for name, (io, fn) in arms.items():
            sm.set_option("io", io)
            torch.cuda.synchronize()
            import time
            time.sleep(0.5)   # start every burst from an idle GPU
            ts, c = timeline(fn, count)
            steady = round(sum(ts[-20:]) / 20, 4)
            out[f"{name}_{n}"] = {"ms": ts, "clocks_end": c, "first5_mean": round(sum(ts[:5]) / 5, 4), "steady_last20_mean": steady}
            print(name, n, "first5", ts[:5], "last5", ts[-5:], "steady", steady, "| clocks", c, flush=True)
        cu.cufftDestroy(h)
    sm.set_option("io", 0)
    json.dump(out, open(out_path, "w"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/burst_timeline.json", int(sys.argv[2]) if len(sys.argv) > 2 else 80)
