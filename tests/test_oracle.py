"""Pins the CPU oracle (oracle/smfft_oracle.c) before anything trusts it.

The reference ships no golden vectors or known-answer tests, so the pins are:
  * an independent FP64 O(N^2) DFT written from the definition (oracle_dft64_direct),
  * numpy's FP64 FFT through the closed forms of SURVEY.md appendix A.1,
  * one-hot known answers (appendix A.1: reorder y[k] = W^{pk}; no-reorder y[k] = W^{brev(p) k}),
  * the exact integer permutation brev_e,
  * fixtures generated on a B200 by the reference's own recompiled kernels (tests/golden/, when present).
Tolerance: relative L2 <= 1e-5 (north_star) -- an fp32 radix-2 FFT sits near 1e-7.
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle_np as O

SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096]
TOL = 1e-5


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("inverse", [False, True])
def test_fp64_direct_dft_matches_numpy(n, inverse):
    x = O.uniform_c64(2, n)
    ref = np.fft.ifft(x.astype(np.complex128), axis=-1) * n if inverse else np.fft.fft(x.astype(np.complex128), axis=-1)
    assert O.rel_l2(O.c_dft64(x, inverse), ref) < 1e-12


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("reorder", [False, True])
def test_ct_restatement_vs_fp64(n, inverse, reorder):
    x = O.uniform_c64(3, n)
    got = O.c_ct_c2c(x, inverse, reorder)
    assert O.rel_l2(got, O.ct_c2c_fp64(x, inverse, reorder)) < TOL
    # and against the from-the-definition DFT, applied to the permuted input for no-reorder
    xin = x if reorder else x[:, O.brev_perm(n)]
    assert O.rel_l2(got, O.c_dft64(xin, inverse)) < TOL


def test_ct_4096_inverse_noreorder_quirk():
    # CT/SM_FFT_parameters.cuh:388: fft_direction = 0 in FFT_4096_inverse_noreorder
    x = O.uniform_c64(1, 4096)
    quirk = O.c_ct_c2c(x, True, False, quirk_4096=True)
    assert O.rel_l2(quirk, O.ct_c2c_fp64(x, False, False)) < TOL
    assert O.rel_l2(quirk, O.ct_c2c_fp64(x, True, False, quirk_4096=True)) < TOL
    assert O.rel_l2(O.c_ct_c2c(x, True, False), O.ct_c2c_fp64(x, True, False)) < TOL


@pytest.mark.parametrize("n", SIZES)
def test_reorder_permutation_exact(n):
    perm = O.c_reorder_index(n)
    assert np.array_equal(perm, O.brev_perm(n))
    assert np.array_equal(perm[perm], np.arange(n))  # involution
    # one-hot known answers pin the permutation through the transform itself
    e = n.bit_length() - 1
    for p in (1, 3, n // 2 + 1, n - 1):
        x = np.zeros((1, n), np.complex64)
        x[0, p] = 1
        k = np.arange(n)
        y1 = O.c_ct_c2c(x, False, True)[0]
        y0 = O.c_ct_c2c(x, False, False)[0]
        bp = int(format(p, f"0{e}b")[::-1], 2)
        assert np.allclose(y1, np.exp(-2j * np.pi * p * k / n), atol=2e-5)
        assert np.allclose(y0, np.exp(-2j * np.pi * bp * k / n), atol=2e-5)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("inverse", [False, True])
def test_stockham_restatement_vs_fp64(n, inverse):
    x = O.uniform_c64(3, n)
    assert O.rel_l2(O.c_stockham_c2c(x, inverse), O.stockham_c2c_fp64(x, inverse)) < TOL


@pytest.mark.parametrize("n", [128, 256, 512, 1024, 2048, 4096, 8192])
def test_r2c_c2r_restatement_vs_fp64(n):
    x = O.uniform_f32(3, n)
    y = O.c_r2c(x)
    assert y.shape == (3, n // 2)
    assert O.rel_l2(y, O.r2c_packed_fp64(x)) < TOL
    # C2R on a random packed half-spectrum with real DC/Nyquist (RC/FFT.c:264-283)
    h = O.uniform_c64(3, n // 2, seed=7)
    assert O.rel_l2(O.c_c2r(h), O.c2r_packed_fp64(h)) < TOL
    # round trip scales by N/2 (appendix A.1)
    assert O.rel_l2(O.c_c2r(y) / (n / 2), x) < TOL


def test_linearity_and_parseval():
    n = 1024
    a, b = O.uniform_c64(2, n, seed=1), O.uniform_c64(2, n, seed=2)
    fa, fb, fab = O.c_ct_c2c(a, False, True), O.c_ct_c2c(b, False, True), O.c_ct_c2c(a + 2 * b, False, True)
    assert O.rel_l2(fab, fa.astype(np.complex128) + 2 * fb.astype(np.complex128)) < TOL
    assert abs(np.sum(np.abs(fa) ** 2) / (n * np.sum(np.abs(a) ** 2)) - 1) < 1e-5
    # forward then inverse = N * identity (both un-normalised)
    assert O.rel_l2(O.c_ct_c2c(fa, True, True) / n, a) < TOL


def test_reference_comparator_restatement():
    # get_error/Compare_data (CT/FFT.c:23-77): abs-value compare, 1e-4 threshold, decade scaling above 10
    a = np.array([[1 + 1j, 100 + 0j, -2 + 0j]], np.complex64)
    assert O.c_ref_compare(a, a) == 0
    b = a.copy(); b[0, 0] += 1e-3
    assert O.c_ref_compare(a, b) == 1
    c = a.copy(); c[0, 1] += 5e-3           # 5e-3 / 10^2 = 5e-5 < 1e-4
    assert O.c_ref_compare(a, c) == 0
    d = a.copy(); d[0, 2] = 2               # sign flip is invisible to the reference's check
    assert O.c_ref_compare(a, d) == 0


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))


@pytest.mark.parametrize("path", GOLDEN or [None])
def test_oracle_vs_reference_golden(path):
    """Fixtures written by tests/golden/make_golden.py on a B200: outputs of the reference's own
    kernels (oracle/_ref, unmodified sources, sm_100a) on seeded inputs."""
    if path is None:
        pytest.skip("no reference fixtures committed yet (generated on the GPU box)")
    g = np.load(path)
    kind = str(g["kind"])
    x = g["input"]
    if kind == "ct":
        got = O.c_ct_c2c(x, bool(g["inverse"]), bool(g["reorder"]), quirk_4096=True)
    elif kind == "stockham":
        got = O.c_stockham_c2c(x, True)
    elif kind == "r2c":
        got = O.c_r2c(x)
    elif kind == "c2r":
        got = O.c_c2r(x)
    else:
        pytest.fail(kind)
    assert O.rel_l2(got, g["output"]) < TOL
