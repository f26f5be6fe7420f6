#!/usr/bin/env python
"""tools/select_probe.py -- what the first-use selection (option "select" = 1) picks on this box at the bench's batch size,
and what that is worth inside the sustained 16-launch step: the step is run alternately with the static table ("select" = 0)
and with the selected instances (A B A B ...), per-launch CUDA events, median per size.  Also runs every register-direct
shape explicitly (io = 4 / 5) in the same step, so the table can be re-derived from one file.

    python tools/select_probe.py [out.json] [steps]
"""
import ctypes
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


_CU = {}


def cufft_step(x, y, rec):
    """cuFFT has no no-reorder mode: every size twice, like bench.py's cuFFT arm"""
    if not _CU:
        lib = ctypes.CDLL("libcufft.so.11")
        _CU["lib"] = lib
        for n in SIZES:
            h = ctypes.c_int(0)
            assert lib.cufftPlan1d(ctypes.byref(h), n, 0x29, PTS // n) == 0
            _CU[n] = h
    lib = _CU["lib"]
    for n in SIZES:
        for reorder in (1, 0):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.cufftExecC2C(_CU[n], ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), -1)
            e1.record()
            rec.setdefault(f"{n}{'r' if reorder else 'n'}", []).append((e0, e1))


def step(x, y, rec):
    for n in SIZES:
        for reorder in (1, 0):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sm.exec_c2c(x, y, n, PTS // n, False, bool(reorder))
            e1.record()
            rec.setdefault(f"{n}{'r' if reorder else 'n'}", []).append((e0, e1))


def main(out_path, steps):
    torch.cuda.set_device(0)
    sm.FFT_init()
    x = torch.rand((PTS, 2), device="cuda")
    y = torch.empty_like(x)
    sm.set_option("select_reset", 1)
    sm.set_option("select", 1)
    step(x, y, {})   # tunes every transform of the step
    torch.cuda.synchronize()
    report = sm.select_report()
    print(report)
    arms = {"static": lambda: (sm.set_option("select", 0), sm.set_option("io", 0)),
            "selected": lambda: (sm.set_option("select", 1), sm.set_option("io", 0)),
            "reg_a": lambda: (sm.set_option("select", 0), sm.set_option("io", 4)),
            "reg_b": lambda: (sm.set_option("select", 0), sm.set_option("io", 5))}
    recs = {k: {} for k in arms}
    recs["cufft"] = {}
    for name, setup in arms.items():   # warm every arm
        setup()
        step(x, y, {})
    cufft_step(x, y, {})
    torch.cuda.synchronize()
    for _ in range(steps):
        for name, setup in arms.items():
            setup()
            step(x, y, recs[name])
        cufft_step(x, y, recs["cufft"])
    torch.cuda.synchronize()
    # the same arms again, each ALONE for `steps` consecutive steps (what bench.py's timed region and baseline arms do)
    alone = {}
    for name in ("static", "cufft", "reg_a"):
        if name != "cufft":
            arms[name]()
        rec = {}
        for i in range(steps + 3):
            (cufft_step if name == "cufft" else step)(x, y, rec if i >= 3 else {})
        torch.cuda.synchronize()
        alone[name] = {k: round(statistics.median(a.elapsed_time(b) for a, b in v), 4) for k, v in rec.items()}
    sm.set_option("select", 0)
    sm.set_option("io", 0)
    out = {"select_report": report.splitlines(), "steps": steps, "per_size_ms": {}, "step_ms": {}}
    for name, rec in recs.items():
        med = {k: round(statistics.median(a.elapsed_time(b) for a, b in v), 4) for k, v in rec.items()}
        out["per_size_ms"][name] = med
        out["step_ms"][name] = round(sum(med.values()), 4)
    out["alone_per_size_ms"] = alone
    out["alone_step_ms"] = {k: round(sum(v.values()), 4) for k, v in alone.items()}
    print("alone", out["alone_step_ms"], {k: {s_: v[s_] for s_ in ("128r", "256r", "512r", "1024r", "4096r")} for k, v in alone.items()})
    print(json.dumps(out["per_size_ms"], indent=0))
    print(out["step_ms"])
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/select_probe.json", int(sys.argv[2]) if len(sys.argv) > 2 else 10)
