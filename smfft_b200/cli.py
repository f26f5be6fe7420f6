"""Command-line mirror of the reference's FFT.exe programs, reproducible and with a real error metric.

    python -m smfft_b200.cli c2c  FFT_size nFFTs nRuns inverse reorder     # SMFFT_CooleyTukey_C2C/FFT.c:84-100
    python -m smfft_b200.cli stockham FFT_size nFFTs nRuns                  # SMFFT_Stockham_C2C/FFT.c:85-97 (inverse)
    python -m smfft_b200.cli r2c  FFT_size nFFTs nRuns                      # SMFFT_Stockham_R2C_C2R/FFT.c:194-206 (R2C then C2R)

Differences from the reference program: the input is seeded (20260101) instead of srand(time(NULL)); the check is a
relative L2 error against an FP64 FFT (numpy) -- plus the reference's own |A|-|B| <= 1e-4 criterion for continuity --
and it also covers the no-reorder transform, which the reference never verifies (CT/FFT.c:161-163).
The FP64 check runs on the host over at most 4096 transforms; the GPU path has no CPU fallback.
"""
from __future__ import annotations

import sys

import numpy as np


def _brev(n):
    e = n.bit_length() - 1
    i = np.arange(n)
    out = np.zeros(n, dtype=np.int64)
    for b in range(e):
        out |= ((i >> b) & 1) << (e - 1 - b)
    return out


def _rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def _ref_pass(a, b, max_error=1e-4):
    """get_error / Compare_data (CT/FFT.c:23-77): |A| vs |B|, decade-scaled above 10, count of offenders."""
    def err(p, q):
        p, q = np.abs(p), np.abs(q)
        small = np.minimum(p, q)
        scale = np.where(small > 10, 10.0 ** np.floor(np.log10(np.maximum(small, 1e-30))), 1.0)
        return np.abs(p - q) / scale
    e = np.maximum(err(a.real, b.real), err(a.imag, b.imag)) if np.iscomplexobj(a) else err(a, b)
    return int(np.count_nonzero(e > max_error))


def main(argv=None):
    import torch

    import smfft_b200 as sm

    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("c2c", "stockham", "r2c") or len(argv) < 4:
        print(__doc__)
        return 1
    kind, n, nffts, nruns = argv[0], int(argv[1]), int(argv[2]), max(1, int(argv[3]))
    inverse = bool(int(argv[4])) if kind == "c2c" and len(argv) > 4 else (kind == "stockham")
    reorder = bool(int(argv[5])) if kind == "c2c" and len(argv) > 5 else True
    rng = np.random.default_rng(20260101)
    ncheck = min(nffts, 4096)
    sm.FFT_init()
    if kind in ("c2c", "stockham"):
        x = rng.random((nffts, n, 2), dtype=np.float32)
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty_like(dx)
        t_ext = sum(sm.FFT_external_benchmark(dx, dy, n, nffts, inverse, reorder) for _ in range(nruns)) / nruns
        y = dy[:ncheck].cpu().numpy().view(np.complex64).reshape(ncheck, n)
        t_mul = -1.0
        if nffts >= 100:
            scratch = torch.empty_like(dx)
            t_mul = sum(sm.FFT_multiple_benchmark(dx, scratch, n, nffts, inverse, reorder) for _ in range(nruns)) / nruns
        xc = x[:ncheck].view(np.complex64).reshape(ncheck, n).astype(np.complex128)
        if not reorder:
            xc = xc[:, _brev(n)]
        want = np.fft.ifft(xc, axis=-1) * n if inverse else np.fft.fft(xc, axis=-1)
        print(f"  SH FFT normal = {t_ext:0.3f} ms; SM FFT multiple times = {t_mul:0.3f} ms")
        rel, bad = _rel_l2(y, want), _ref_pass(y, want.astype(np.complex64))
        print(f"  FFT size: {n}; nFFTs: {nffts}; inverse={int(inverse)} reorder={int(reorder)}; "
              f"{nffts * n * 16 / t_ext / 1e6:0.1f} GB/s; relative L2 vs FP64 = {rel:.3e}; reference-criterion errors = {bad}")
        ok = rel < 1e-5
    else:
        x = rng.random((nffts, n), dtype=np.float32)
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty((nffts, n // 2, 2), dtype=torch.float32, device="cuda")
        dz = torch.empty_like(dx)
        t_f = sum(sm.R2C_C2R_external_benchmark(dx, dy, n, nffts, 0) for _ in range(nruns)) / nruns
        t_i = sum(sm.R2C_C2R_external_benchmark(dy, dz, n, nffts, 1) for _ in range(nruns)) / nruns
        y = dy[:ncheck].cpu().numpy().view(np.complex64).reshape(ncheck, n // 2)
        full = np.fft.rfft(x[:ncheck].astype(np.float64), axis=-1)
        want = full[:, : n // 2].copy()
        want[:, 0] = full[:, 0].real + 1j * full[:, n // 2].real
        rel_f = _rel_l2(y, want)
        rel_i = _rel_l2(dz[:ncheck].cpu().numpy() / (n / 2), x[:ncheck])
        print(f"  R2C = {t_f:0.3f} ms ({nffts * n * 8 / t_f / 1e6:0.1f} GB/s); C2R = {t_i:0.3f} ms; "
              f"relative L2: R2C vs FP64 = {rel_f:.3e}, C2R(R2C(x))/(N/2) vs x = {rel_i:.3e}")
        ok = rel_f < 1e-5 and rel_i < 1e-5
    print("  FFT test: " + ("PASSED" if ok else "FAILED"))
    return 0 if ok else 2


if __name__ == "__main__":
    sys.exit(main())
