#!/usr/bin/env python
"""tools/report.py -- full comparison report on one B200 (measurement tool; writes JSON).

For every size: ours (TMA staging; table and MUFU twiddles) vs the reference's own kernels rebuilt for
sm_100a (oracle/_ref) vs cuFFT, for FFT_external (HBM-bound) and FFT_multiple (compute-bound); Stockham
and R2C/C2R likewise; plus the accuracy of each against an FP64 DFT (SURVEY.md 7.3 first on-box task).
Same protocol for every row: cudaEvent timing on one stream, data resident, 3 warm-ups, `reps` timed
runs, median and min.
"""
import ctypes
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smfft_b200 as sm  # noqa: E402
from oracle import oracle_np as O  # noqa: E402
from tests import refkernels as R  # noqa: E402

SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096]
PTS = 1 << 29


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return {"ms": round(statistics.median(ts), 4), "ms_min": round(min(ts), 4)}


class CuFFT:
    C2C, R2C, C2R = 0x29, 0x2A, 0x2C

    def __init__(self):
        self.lib = None
        for name in ("libcufft.so.11", "libcufft.so"):
            try:
                self.lib = ctypes.CDLL(name)
                break
            except OSError:
                continue
        self.plans = {}

    def plan(self, n, kind, batch):
        key = (n, kind, batch)
        if key not in self.plans:
            h = ctypes.c_int(0)
            rc = self.lib.cufftPlan1d(ctypes.byref(h), n, kind, batch)
            if rc != 0:
                raise RuntimeError(f"cufftPlan1d rc={rc}")
            self.plans[key] = h
        return self.plans[key]

    def c2c(self, x, y, n, batch, inverse):
        p = self.plan(n, self.C2C, batch)
        return lambda: self.lib.cufftExecC2C(p, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), 1 if inverse else -1)

    def r2c(self, x, y, n, batch):
        p = self.plan(n, self.R2C, batch)
        return lambda: self.lib.cufftExecR2C(p, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()))

    def c2r(self, x, y, n, batch):
        p = self.plan(n, self.C2R, batch)
        return lambda: self.lib.cufftExecC2R(p, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()))

    def free(self):
        for h in self.plans.values():
            self.lib.cufftDestroy(h)
        self.plans = {}


def main(out_path, reps=10):
    torch.cuda.set_device(0)
    sm.FFT_init()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260101)
    x = torch.rand((PTS, 2), device="cuda", generator=gen)
    y = torch.empty((PTS + PTS // 16 + 8192, 2), device="cuda")  # cuFFT R2C writes N/2+1 bins per transform (33/32 at N=64)
    rep = {"device": torch.cuda.get_device_name(0), "points": PTS, "reps": reps, "ct_external": {}, "ct_multiple": {},
           "stockham": {}, "r2c_c2r": {}, "accuracy": {}}
    cu = CuFFT()
    have_ref = R.available()
    gb = PTS * 16 / 1e6

    for n in SIZES:
        nf = PTS // n
        for reorder in (1, 0):
            key = f"{n}{'r' if reorder else 'n'}"
            row = {}
            for tw, name in ((0, "ours_lut"), (1, "ours_mufu")):
                sm.set_option("twiddle", tw)
                row[name] = timeit(lambda: sm.exec_c2c(x, y, n, nf, False, bool(reorder)), reps)
            sm.set_option("twiddle", 0)
            row["ours_inverse"] = timeit(lambda: sm.exec_c2c(x, y, n, nf, True, bool(reorder)), reps)
            if have_ref:
                row["reference_sm100a"] = timeit(lambda: R.ct_external(x, y, n, nf, False, reorder), reps)
            if reorder and cu.lib is not None:
                try:
                    row["cufft"] = timeit(cu.c2c(x, y, n, nf, False), reps)
                    cu.free()
                except Exception as ex:
                    row["cufft"] = {"error": str(ex)}
            for v in row.values():
                if "ms" in v:
                    v["GBps"] = round(gb / v["ms"], 1)
            rep["ct_external"][key] = row
            # FFT_multiple: nFFTs/100 transforms' worth of data, 100 in-place reps (compute-bound)
            mrow = {}
            flops = (nf // 100) * 100 * 5.0 * n * np.log2(n)
            for tw, name in ((0, "ours_lut"), (1, "ours_mufu")):
                sm.set_option("twiddle", tw)
                ts = [sm.FFT_multiple_benchmark(x, y, n, nf, False, bool(reorder)) for _ in range(reps + 2)][2:]
                mrow[name] = {"ms": round(statistics.median(ts), 4), "ms_min": round(min(ts), 4)}
            sm.set_option("twiddle", 0)
            if have_ref:
                ts = [R.ct_multiple(x, y, n, nf, False, reorder) for _ in range(reps + 2)][2:]
                mrow["reference_sm100a"] = {"ms": round(statistics.median(ts), 4), "ms_min": round(min(ts), 4)}
            for v in mrow.values():
                v["TFLOPs"] = round(flops / v["ms"] / 1e9, 2)
            rep["ct_multiple"][key] = mrow
        print("ct", n, json.dumps(rep["ct_external"][f"{n}r"]), json.dumps(rep["ct_multiple"][f"{n}r"]), flush=True)

    for n in (256, 512, 1024, 2048, 4096):
        nf = PTS // n
        row = {"ours_inverse": timeit(lambda: sm.exec_c2c(x, y, n, nf, True, True), reps),
               "ours_forward": timeit(lambda: sm.exec_c2c(x, y, n, nf, False, True), reps)}
        ts = [sm.Stockham_multiple_benchmark(x, y, n, nf, True) for _ in range(reps + 2)][2:]
        row["ours_multiple"] = {"ms": round(statistics.median(ts), 4)}
        if have_ref:
            row["reference_sm100a"] = timeit(lambda: R.st_external(x, y, n, nf), reps)
            ts = [R.st_multiple(x, y, n, nf) for _ in range(reps + 2)][2:]
            row["reference_multiple"] = {"ms": round(statistics.median(ts), 4)}
        rep["stockham"][str(n)] = row
    print("stockham", json.dumps(rep["stockham"]), flush=True)

    xr = x.view(-1)  # 2^30 reals = 4 GiB
    for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192):
        nf = (2 * PTS) // n
        row = {"r2c_ours": timeit(lambda: sm.exec_r2c_c2r(xr, y, n, nf, 0), reps),
               "c2r_ours": timeit(lambda: sm.exec_r2c_c2r(xr, y, n, nf, 1), reps)}
        if n >= 128:
            ts = [sm.R2C_multiple_benchmark(xr, y, n, nf) for _ in range(reps + 2)][2:]
            row["r2c_multiple_ours"] = {"ms": round(statistics.median(ts), 4)}
        if have_ref and 512 <= n <= 4096:
            row["r2c_reference_sm100a"] = timeit(lambda: R.rc_external(xr, y, n, nf, 0), reps)
            row["c2r_reference_sm100a"] = timeit(lambda: R.rc_external(xr, y, n, nf, 1), reps)
        if cu.lib is not None:
            try:
                row["r2c_cufft"] = timeit(cu.r2c(xr, y, n, nf), reps)
                cu.free()
                row["c2r_cufft"] = timeit(cu.c2r(y, xr, n, nf), reps)  # clobbers xr (timing only)
                cu.free()
                xr.copy_(torch.rand(xr.shape, device="cuda", generator=gen))
            except Exception as ex:
                row["cufft"] = {"error": str(ex)}
        for v in row.values():
            if "ms" in v:
                v["GBps"] = round(gb / v["ms"], 1)
        rep["r2c_c2r"][str(n)] = row
        print("r2c", n, json.dumps(row), flush=True)

    # accuracy vs FP64 on 64 transforms per size (relative L2)
    for n in SIZES:
        xs = O.uniform_c64(64, n)
        dx = torch.from_numpy(xs.view(np.float32).reshape(64, n, 2)).cuda()
        dy = torch.zeros_like(dx)
        want = O.ct_c2c_fp64(xs, False, True)
        acc = {}
        for tw, name in ((0, "ours_lut"), (1, "ours_mufu")):
            sm.set_option("twiddle", tw)
            sm.exec_c2c(dx, dy, n, 64, False, True)
            torch.cuda.synchronize()
            acc[name] = O.rel_l2(dy.cpu().numpy().view(np.complex64).reshape(64, n), want)
        sm.set_option("twiddle", 0)
        if have_ref:
            R.ct_external(dx, dy, n, 64, False, True)
            torch.cuda.synchronize()
            ref = dy.cpu().numpy().view(np.complex64).reshape(64, n)
            acc["reference_sm100a"] = O.rel_l2(ref, want)
            sm.exec_c2c(dx, dy, n, 64, False, True)
            torch.cuda.synchronize()
            acc["ours_vs_reference"] = O.rel_l2(dy.cpu().numpy().view(np.complex64).reshape(64, n), ref)
        if cu.lib is not None:
            cu.c2c(dx, dy, n, 64, False)()
            torch.cuda.synchronize()
            acc["cufft"] = O.rel_l2(dy.cpu().numpy().view(np.complex64).reshape(64, n), want)
            cu.free()
        acc["cpu_oracle_f32"] = O.rel_l2(O.c_ct_c2c(xs, False, True), want)
        rep["accuracy"][str(n)] = {k: float(f"{v:.3e}") for k, v in acc.items()}
    print("accuracy", json.dumps(rep["accuracy"]), flush=True)
    json.dump(rep, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/report.json", int(sys.argv[2]) if len(sys.argv) > 2 else 10)
