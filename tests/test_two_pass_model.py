"""CPU model of the multi-pass factorisations that smfft_b200/csrc/big_fft.cu runs: two passes for 2^15 .. 2^20 points
(n = n1 + N1 n2, k = N2 k1 + k2), three for 2^21 .. 2^24 (n = n1 + N1 n2 + N1 N2 n3, k = k3 + N3 k2 + N2 N3 k1) -- the same index
conventions, the same split of the sizes, the same three-level FP64-rounded twiddle table (W_M^j, W_M^(512 j), W_M^(2^18 j))
evaluated in float32 -- against numpy's FP64 FFT.  Pins the algebra and the table layout without a GPU; the kernels
themselves are checked by tests/test_gpu_parity.py::test_two_pass_transforms."""
import numpy as np
import pytest


def split(e):
    """log2 of (N2 = strided pass A, N1 = contiguous pass B), as big_fft.cu: two passes up to 2^20 points"""
    l2 = 10 if e == 20 else 9 if e >= 18 else 8
    return l2, e - l2


def three_level(m, p):
    """W_M^p = lo[p & 511] mid[(p >> 9) & 511] hi[p >> 18], every table entry rounded from FP64 to float32"""
    j = np.arange(512, dtype=np.int64)
    tab = [np.exp(-2j * np.pi * ((j << sh) % m) / m).astype(np.complex64) for sh in (0, 9, 18)]
    return (tab[0][p & 511] * tab[1][(p >> 9) & 511]).astype(np.complex64) * tab[2][p >> 18]


def two_pass(x, inverse):
    n = x.shape[-1]
    e = n.bit_length() - 1
    l2, l1 = split(e)
    n2, n1 = 1 << l2, 1 << l1
    a = x.reshape(-1, n2, n1)                                    # [fft][n2][n1]: n = n1 + N1 n2
    fa = (np.fft.ifft(a, axis=1) * n2 if inverse else np.fft.fft(a, axis=1)).astype(np.complex64)   # pass A: over n2 -> [fft][k2][n1]
    p = np.arange(n2, dtype=np.int64)[:, None] * np.arange(n1, dtype=np.int64)[None, :]          # n1 k2 < N
    assert p.max() < n
    w = three_level(n, p)                                        # the kernel's W(p)
    if inverse:
        w = np.conj(w)
    b = fa * w[None]
    fb = (np.fft.ifft(b, axis=2) * n1 if inverse else np.fft.fft(b, axis=2)).astype(np.complex64)   # pass B: over n1 -> [fft][k2][k1]
    return np.transpose(fb, (0, 2, 1)).reshape(-1, n)            # X[N2 k1 + k2]: rows k1, columns k2


@pytest.mark.parametrize("e", [15, 16, 17, 18, 19, 20])
@pytest.mark.parametrize("inverse", [False, True])
def test_factorisation_and_twiddle_table(e, inverse):
    n = 1 << e
    rng = np.random.default_rng(e)
    x = rng.random((2, n, 2), dtype=np.float32).view(np.complex64).reshape(2, n)
    want = np.fft.ifft(x.astype(np.complex128), axis=-1) * n if inverse else np.fft.fft(x.astype(np.complex128), axis=-1)
    got = two_pass(x, inverse)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel < 1e-6, rel


def split3(e):
    """log2 of (N3, N2, N1) for the three-pass sizes, as big_fft.cu"""
    l1 = (e + 2) // 3
    l2 = (e - l1 + 1) // 2
    return e - l1 - l2, l2, l1


@pytest.mark.parametrize("e", [21, 22])
def test_three_pass_factorisation(e):
    n = 1 << e
    l3, l2, l1 = split3(e)
    n3, n2, n1 = 1 << l3, 1 << l2, 1 << l1
    assert l1 + l2 + l3 == e and all(6 <= v <= 8 for v in (l1, l2, l3))
    rng = np.random.default_rng(e)
    x = rng.random((1, n, 2), dtype=np.float32).view(np.complex64).reshape(1, n)
    a = x.reshape(1, n3, n2, n1)                                              # [fft][n3][n2][n1]
    a = np.fft.fft(a, axis=1).astype(np.complex64)                            # pass 1: over n3 -> [k3][n2][n1]
    k3 = np.arange(n3, dtype=np.int64)[:, None, None]
    i2 = np.arange(n2, dtype=np.int64)[None, :, None]
    i1 = np.arange(n1, dtype=np.int64)[None, None, :]
    a = a * three_level(n2 * n3, (i2 * k3) + 0 * i1)[None]                    # W_(N2 N3)^(n2 k3)
    a = np.fft.fft(a, axis=2).astype(np.complex64)                            # pass 2: over n2 -> [k3][k2][n1]
    p = i1 * (k3 + n3 * i2)                                                   # n1 (k3 + N3 k2) < N
    assert p.max() < n
    a = a * three_level(n, p)[None]
    a = np.fft.fft(a, axis=3).astype(np.complex64)                            # pass 3: over n1 -> [k3][k2][k1]
    got = np.transpose(a, (0, 3, 2, 1)).reshape(1, n)                         # X[k3 + N3 k2 + N2 N3 k1]
    want = np.fft.fft(x.astype(np.complex128), axis=-1)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-6


def test_split_covers_the_range():
    for e in range(21, 25):
        assert sum(split3(e)) == e and all(6 <= v <= 8 for v in split3(e))  # block transforms of 64 .. 256 points
    for e in range(15, 21):
        l2, l1 = split(e)
        assert l1 + l2 == e and 7 <= l1 <= 10 and 8 <= l2 <= 10   # block transforms of 128 .. 1024 points (1024: 8 per tile, else 16)
