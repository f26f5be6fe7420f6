"""The two in-place stage formulas include/smfft/detail/warp_fft.cuh is built on, restated in numpy at array level and
checked against an FP64 FFT: DIF (natural in, position p ends with X[brev(p)], twiddle AFTER the butterfly, output digit
stored bit-swapped) and DIT (computes DFT(x o brev) in natural order, twiddle BEFORE the butterfly on the bit-swapped
position digit).  Radix 4 with one radix-2 stage for odd log2 N, exactly the stage lists of the warp plans."""
import numpy as np
import pytest


def brev(p, bits):
    r = 0
    for i in range(bits):
        r |= ((p >> i) & 1) << (bits - 1 - i)
    return r


def dif_inplace(x, s, sign=-1):
    a_ = x.astype(np.complex128).copy()
    m_ = 1 << s
    h = s - 1
    while h >= 0:
        bits = (h, h - 1) if h >= 1 else (h,)
        r = 1 << len(bits)
        lowmask = (1 << bits[-1]) - 1
        b_ = a_.copy()
        for p in range(m_):
            k = (((p >> bits[0]) & 1) | (((p >> bits[1]) & 1) << 1)) if r == 4 else (p >> bits[0]) & 1   # (P_hi, P_lo) = (k0, k1)
            base = p
            for b in bits:
                base &= ~(1 << b)
            acc = 0
            for a in range(r):
                q = base | ((((a >> 1) & 1) << bits[0] | (a & 1) << bits[1]) if r == 4 else a << bits[0])
                acc += a_[q] * np.exp(sign * 2j * np.pi * a * k / r)
            b_[p] = acc * np.exp(sign * 2j * np.pi * k * (p & lowmask) / (1 << (h + 1)))
        a_ = b_
        h -= len(bits)
    return a_


def dit_inplace(y, s, sign=-1):
    a_ = y.astype(np.complex128).copy()
    m_ = 1 << s
    lo = 0
    while lo < s:
        bits = (lo + 1, lo) if lo + 1 < s else (lo,)
        r = 1 << len(bits)
        h = bits[0]
        b_ = a_.copy()
        for p in range(m_):
            kk = (((p >> bits[1]) & 1) | (((p >> bits[0]) & 1) << 1)) if r == 4 else (p >> bits[0]) & 1
            base = p
            for b in bits:
                base &= ~(1 << b)
            klow = p & ((1 << lo) - 1)
            acc = 0
            for d in range(r):
                if r == 4:
                    q = base | ((d >> 1) & 1) << bits[0] | (d & 1) << bits[1]
                    a = ((d & 1) << 1) | (d >> 1)      # input index = bit-swapped position digit
                else:
                    q, a = base | d << bits[0], d
                acc += a_[q] * np.exp(sign * 2j * np.pi * a * klow / (1 << (h + 1))) * np.exp(sign * 2j * np.pi * a * kk / r)
            b_[p] = acc
        a_ = b_
        lo += len(bits)
    return a_


@pytest.mark.parametrize("s", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("sign", [-1, 1])
def test_inplace_stage_formulas(s, sign):
    m = 1 << s
    rng = np.random.default_rng(s)
    x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    full = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * m
    perm = np.array([brev(p, s) for p in range(m)])
    assert np.abs(dif_inplace(x, s, sign) - full[perm]).max() < 1e-12 * m
    want = np.fft.fft(x[perm]) if sign < 0 else np.fft.ifft(x[perm]) * m
    assert np.abs(dit_inplace(x, s, sign) - want).max() < 1e-12 * m
