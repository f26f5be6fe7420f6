"""FP64 closed-form oracle for the SMFFT hot path, plus the ctypes loader for smfft_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
Nothing under smfft_b200/ imports this module; the product path has no CPU fallback.

Closed forms follow SURVEY.md appendix A.1 (verified there against a lane-exact emulation of the
reference kernels) and cite the reference code they describe:
  CT = SMFFT_CooleyTukey_C2C/FFT-GPU-32bit.cu, RC = SMFFT_Stockham_R2C_C2R/FFT-GPU-32bit-Stockham.cu
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20260101


# ----------------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------------
def uniform_c64(nffts: int, n: int, seed: int = SEED) -> np.ndarray:
    """re, im i.i.d. uniform [0,1): the reference's distribution (CT/FFT.c:139-143), fixed seed."""
    rng = np.random.default_rng(seed)
    a = rng.random((nffts, n, 2), dtype=np.float32)
    return a.view(np.complex64).reshape(nffts, n)


def uniform_f32(nffts: int, n: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.random((nffts, n), dtype=np.float32)


def brev_perm(n: int) -> np.ndarray:
    """brev_e(i) for i in [0, n): the net slot->source map of reorder_32..4096 (CT:54-329)."""
    e = n.bit_length() - 1
    idx = np.arange(n, dtype=np.int64)
    out = np.zeros(n, dtype=np.int64)
    for b in range(e):
        out |= ((idx >> b) & 1) << (e - 1 - b)
    return out


# ----------------------------------------------------------------------------------------------
# FP64 closed forms (appendix A.1)
# ----------------------------------------------------------------------------------------------
def ct_c2c_fp64(x: np.ndarray, inverse: bool, reorder: bool, quirk_4096: bool = False) -> np.ndarray:
    """do_SMFFT_CT_DIT<P> (CT:334-532): reorder=1 -> natural-order un-normalised DFT;
    reorder=0 -> DFT of the bit-reversed INPUT, natural order of k."""
    x = np.asarray(x, dtype=np.complex128)
    n = x.shape[-1]
    if quirk_4096 and n == 4096 and inverse and not reorder:
        inverse = False  # CT/SM_FFT_parameters.cuh:388
    if not reorder:
        x = x[..., brev_perm(n)]
    return np.fft.ifft(x, axis=-1) * n if inverse else np.fft.fft(x, axis=-1)


def stockham_c2c_fp64(x: np.ndarray, inverse: bool) -> np.ndarray:
    """do_FFT_Stockham_C2C<P,Dir> (RC:106-266) / do_FFT_Stockham_mk6 (ST:97-240, inverse only)."""
    return ct_c2c_fp64(x, inverse, True)


def r2c_packed_fp64(x: np.ndarray) -> np.ndarray:
    """do_FFT_Stockham_R2C_C2R<P,FFT_forward> (RC:269-344): M = N/2 packed bins per transform,
    y[0] = (X[0].re, X[M].re), y[k] = X[k] for k = 1..M-1."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    m = n // 2
    full = np.fft.rfft(x, axis=-1)
    out = full[..., :m].copy()
    out[..., 0] = full[..., 0].real + 1j * full[..., m].real
    return out


def c2r_packed_fp64(y: np.ndarray) -> np.ndarray:
    """do_FFT_Stockham_R2C_C2R<P,FFT_inverse>: packed M bins -> N reals = (N/2) * irfft(Y)
    (= cuFFT C2R / 2; RC/FFT.c:170-171)."""
    y = np.asarray(y, dtype=np.complex128)
    m = y.shape[-1]
    n = 2 * m
    full = np.zeros(y.shape[:-1] + (m + 1,), dtype=np.complex128)
    full[..., :m] = y
    full[..., 0] = y[..., 0].real
    full[..., m] = y[..., 0].imag
    return np.fft.irfft(full, n=n, axis=-1) * (n / 2)


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """relative L2 error ||a-b|| / ||b|| over the whole batch, in fp64."""
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


# ----------------------------------------------------------------------------------------------
# C restatement loader
# ----------------------------------------------------------------------------------------------
_LIB = None


def build_c_oracle() -> str:
    so = os.path.join(HERE, "_build", "libsmfft_oracle.so")
    src = os.path.join(HERE, "smfft_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return so


def c_oracle():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        P, I, LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        lib.oracle_dft64_direct.argtypes = [P, P, I, LL, I]
        lib.oracle_ct_c2c_f32.argtypes = [P, P, I, LL, I, I, I]
        lib.oracle_stockham_c2c_f32.argtypes = [P, P, I, LL, I]
        lib.oracle_r2c_f32.argtypes = [P, P, I, LL]
        lib.oracle_c2r_f32.argtypes = [P, P, I, LL]
        lib.oracle_ct_reorder_index.argtypes = [P, I]
        lib.oracle_ref_compare.argtypes = [P, P, LL, ctypes.c_float]
        lib.oracle_ref_compare.restype = LL
        lib.oracle_num_threads.restype = I
        lib.oracle_set_num_threads.argtypes = [I]
        for f in ("oracle_dft64_direct", "oracle_ct_c2c_f32", "oracle_stockham_c2c_f32", "oracle_r2c_f32",
                  "oracle_c2r_f32", "oracle_ct_reorder_index", "oracle_set_num_threads"):
            getattr(lib, f).restype = None
        _LIB = lib
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def c_dft64(x: np.ndarray, inverse: bool) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty(x.shape, dtype=np.complex128)
    c_oracle().oracle_dft64_direct(_p(x), _p(out), x.shape[-1], x.size // x.shape[-1], int(inverse))
    return out


def c_ct_c2c(x: np.ndarray, inverse: bool, reorder: bool, quirk_4096: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    c_oracle().oracle_ct_c2c_f32(_p(x), _p(out), x.shape[-1], x.size // x.shape[-1], int(inverse), int(reorder),
                                 int(quirk_4096))
    return out


def c_stockham_c2c(x: np.ndarray, inverse: bool) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    c_oracle().oracle_stockham_c2c_f32(_p(x), _p(out), x.shape[-1], x.size // x.shape[-1], int(inverse))
    return out


def c_r2c(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[-1]
    out = np.empty(x.shape[:-1] + (n // 2,), dtype=np.complex64)
    c_oracle().oracle_r2c_f32(_p(x), _p(out), n, x.size // n)
    return out


def c_c2r(y: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.complex64)
    m = y.shape[-1]
    out = np.empty(y.shape[:-1] + (2 * m,), dtype=np.float32)
    c_oracle().oracle_c2r_f32(_p(y), _p(out), 2 * m, y.size // m)
    return out


def c_reorder_index(n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int32)
    c_oracle().oracle_ct_reorder_index(_p(out), n)
    return out


def c_ref_compare(a: np.ndarray, b: np.ndarray, max_error: float = 1e-4) -> int:
    a = np.ascontiguousarray(a, dtype=np.complex64)
    b = np.ascontiguousarray(b, dtype=np.complex64)
    return int(c_oracle().oracle_ref_compare(_p(a), _p(b), a.size, max_error))
