// smfft/detail/twiddle.cuh -- run-time twiddles W_n^m = exp(s 2 pi i m / n).
//
// Two sources, selected per kernel instance by measurement (north_star item 3):
//   TW_LUT  : one accurate base twiddle per pass from a compact per-kernel table in SHARED memory
//             (filled once per persistent CTA from the FP64-rounded global table W_TBL^j, direction
//             already applied), higher powers by complex multiplication in registers.  A lookup is
//             one conflict-free LDS.64; nothing on the critical path depends on L1/L2 residency;
//   TW_MUFU : __sincosf per twiddle (what the reference does on every stage,
//             CT/FFT-GPU-32bit.cu:18-28), argument reduced to [-pi, pi) in integers first.
#pragma once
#include "complex.cuh"

namespace smfft {

enum { TW_LUT = 0, TW_MUFU = 1 };

// table length: W_16384^j serves every C2C modulus up to 16384 and the R2C/C2R pair pass of real N = 8192
constexpr int kTwiddleTableLog2 = 14;
constexpr int kTwiddleTableSize = 1 << kTwiddleTableLog2;

namespace detail {

// W_WN^m from the table; m already reduced to [0, WN)
template <int DIR, int WN>
SMFFT_DEV float2 tw_lut(const float2* __restrict__ tbl, int m)
{
    static_assert(WN <= kTwiddleTableSize, "modulus exceeds the twiddle table");
    float2 w = plat::ldg_ro(tbl + m * (kTwiddleTableSize / WN));
    if (DIR) w.y = -w.y;  // table holds the forward sign
    return w;
}

// W_WN^m by MUFU; m reduced to [-WN/2, WN/2) so the argument stays inside [-pi, pi)
template <int DIR, int WN>
SMFFT_DEV float2 tw_mufu(int m)
{
    m &= (WN - 1);
    if (m >= WN / 2) m -= WN;
    constexpr float kStep = (DIR ? 6.283185307179586f : -6.283185307179586f) / (float)WN;
    float2 w;
    plat::fast_sincos(kStep * (float)m, &w.y, &w.x);
    return w;
}

// pw[q] = W_WN^{k q} for q = 1 .. RAD-1 (pw[0] is not written).
// TW_LUT: `pass_tbl` is this pass's compact shared-memory table, pass_tbl[k] = W_WN^k.
template <int DIR, int TW, int WN, int RAD, class V>
SMFFT_DEV void make_twiddle_powers(V (&pw)[RAD], int k, const float2* pass_tbl)
{
    if constexpr (TW == TW_LUT) {
        pw[1] = from_scalar<V>(plat::lds64(pass_tbl + k));
        static_for<RAD>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            if constexpr (q >= 2) {
                if constexpr ((q & 1) == 0)
                    pw[q] = csqr(pw[q / 2]);
                else
                    pw[q] = cmul(pw[q / 2], pw[q - q / 2]);
            }
        });
    } else {
        static_for<RAD>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            if constexpr (q >= 1) pw[q] = from_scalar<V>(tw_mufu<DIR, WN>(k * q));
        });
    }
}

}  // namespace detail
}  // namespace smfft
