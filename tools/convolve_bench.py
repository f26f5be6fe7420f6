#!/usr/bin/env python
"""tools/convolve_bench.py -- the fused-convolution use case (load -> forward FFT -> pointwise multiply -> inverse FFT ->
store in ONE kernel; README.md:2, 10-14 of the reference), four ways on the same device buffers:

  reference_device_fn   the user kernel on the REFERENCE's do_SMFFT_CT_DIT (oracle/_ref/libsmfft_ref_conv.so)
  compat_device_fn      the same user kernel on include/smfft/compat.cuh (drop-in: same names, same contract)
  native_mufu / _lut    smfft::block_convolve on the native primitive include/smfft/device.cuh (16 points per thread,
                        spectrum multiplied in registers, shared memory only for the exchanges)
  unfused_native        three launches: library forward FFT, torch complex multiply, library inverse FFT
Interleaved timing, CUDA events per launch, median (min).  GB/s = 16 B per point / time (x read + y written).

    python tools/convolve_bench.py [out.json] [reps]
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
from tests.compat.build_compat import build as build_compat  # noqa: E402

PTS = 1 << 28


def main(out_path, reps):
    torch.cuda.set_device(0)
    sm.FFT_init()
    lib = ctypes.CDLL(build_compat())
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.compat_user_convolve.argtypes = [P, P, P, I, I]
    lib.native_convolve_launch.argtypes = [P, P, P, I, I, I, P]
    ref = None
    so = os.path.join(ROOT, "oracle", "_ref", "libsmfft_ref_conv.so")
    if os.path.exists(so):
        ref = ctypes.CDLL(so)
        ref.ref_user_convolve_launch.argtypes = [P, P, P, I, I]
    tw = sm.twiddle_table()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260103)
    x = torch.rand((PTS, 2), device="cuda", generator=gen)
    y = torch.empty_like(x)
    tmp = torch.empty_like(x)
    rep = {"device": torch.cuda.get_device_name(0), "points": PTS, "reps": reps, "sizes": {}}
    for n in (256, 1024, 4096):
        nf = PTS // n
        H = (torch.randn((n, 2), device="cuda", generator=gen) * 0.5).contiguous()
        Hc = torch.view_as_complex(H)
        arms = {}
        if ref is not None:
            arms["reference_device_fn"] = lambda: ref.ref_user_convolve_launch(x.data_ptr(), H.data_ptr(), y.data_ptr(), n, nf)
        arms["compat_device_fn"] = lambda: lib.compat_user_convolve(x.data_ptr(), H.data_ptr(), y.data_ptr(), n, nf)
        arms["native_mufu"] = lambda: lib.native_convolve_launch(x.data_ptr(), H.data_ptr(), y.data_ptr(), n, nf, 0, tw)
        arms["native_lut"] = lambda: lib.native_convolve_launch(x.data_ptr(), H.data_ptr(), y.data_ptr(), n, nf, 1, tw)

        def unfused():
            sm.exec_c2c(x, tmp, n, nf, False, True)
            t = torch.view_as_complex(tmp).view(nf, n)
            t.mul_(Hc).mul_(1.0 / n)
            sm.exec_c2c(tmp, y, n, nf, True, True)

        arms["unfused_native"] = unfused
        # accuracy of each arm on the first 64 transforms vs FP64
        xs = torch.view_as_complex(x[: 64 * n]).view(64, n).cpu().numpy().astype(np.complex128)
        want = np.fft.ifft(np.fft.fft(xs, axis=-1) * Hc.cpu().numpy().astype(np.complex128), axis=-1)
        acc = {}
        for name, fn in arms.items():
            y.zero_()
            fn()
            torch.cuda.synchronize()
            got = torch.view_as_complex(y[: 64 * n]).view(64, n).cpu().numpy()
            acc[name] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        for fn in arms.values():
            for _ in range(2):
                fn()
        torch.cuda.synchronize()
        ts = {k: [] for k in arms}
        for _ in range(reps):
            for name, fn in arms.items():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts[name].append(e0.elapsed_time(e1))
        row = {}
        for name, v in ts.items():
            ms = statistics.median(v)
            row[name] = {"ms": round(ms, 4), "ms_min": round(min(v), 4), "GBps": round(PTS * 16 / ms / 1e6, 1), "rel_l2_vs_fp64": float(f"{acc[name]:.3e}")}
        if "reference_device_fn" in row:
            for name in row:
                row[name]["speedup_vs_reference_device_fn"] = round(row["reference_device_fn"]["ms"] / row[name]["ms"], 3)
        rep["sizes"][str(n)] = row
        print(n, json.dumps(row), flush=True)
    json.dump(rep, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/convolve_bench.json", int(sys.argv[2]) if len(sys.argv) > 2 else 7)
