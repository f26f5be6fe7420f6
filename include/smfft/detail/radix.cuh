// smfft/detail/radix.cuh -- register-resident radix-2/4/8/16/32 DFTs.
//
// A thread keeps RTOT complex values in registers.  dft_regs<DIR, LEN, OFF, STRIDE>() replaces the
// LEN values v[OFF + i*STRIDE] (i = 0..LEN-1) by their LEN-point DFT, natural order in and out.
// The network is a decimation-in-frequency nest with compile-time twiddles; the trailing bit
// reversal is register renaming (every index is a compile-time constant after unrolling).
// PACK = 1 (V = float2): the additions and subtractions of the network are packed FADD2 on (re, im) -- fewer issue
// slots for the issue-bound kernels (profiles/r01_ab_packed_addsub_z.json).
// V is float2 (one transform) or cpair (two transforms in the packed f32x2 lanes, complex.cuh).
// This replaces the reference's one-radix-2-stage-per-shuffle/per-barrier schedule
// (CT/FFT-GPU-32bit.cu:363-531, ST/...:97-240) with log2(LEN) stages per register pass.
#pragma once
#include "complex.cuh"

namespace smfft {
namespace detail {

template <int PACK, int DIR, int LEN, int OFF, int STRIDE, int RTOT, class V>
SMFFT_DEV void dif_net(V (&v)[RTOT])
{
    if constexpr (LEN > 1) {
        constexpr int H = LEN / 2;
        static_for<H>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const V a = v[OFF + i * STRIDE];
            const V b = v[OFF + (i + H) * STRIDE];
            v[OFF + i * STRIDE] = cadd_p<PACK>(a, b);
            v[OFF + (i + H) * STRIDE] = mul_wconst<DIR, i, LEN>(csub_p<PACK>(a, b));
        });
        dif_net<PACK, DIR, H, OFF, STRIDE, RTOT>(v);
        dif_net<PACK, DIR, H, OFF + H * STRIDE, STRIDE, RTOT>(v);
    }
}

template <int PACK, int DIR, int LEN, int OFF, int STRIDE, int RTOT, class V>
SMFFT_DEV void dft_regs(V (&v)[RTOT])
{
    static_assert(OFF + (LEN - 1) * STRIDE < RTOT, "register group out of range");
    dif_net<PACK, DIR, LEN, OFF, STRIDE, RTOT>(v);
    if constexpr (LEN > 2) {
        constexpr int LG = ilog2_c(LEN);
        V t[LEN];
        static_for<LEN>([&](auto I) {
            constexpr int i = decltype(I)::value;
            t[i] = v[OFF + brev_c(i, LG) * STRIDE];
        });
        static_for<LEN>([&](auto I) {
            constexpr int i = decltype(I)::value;
            v[OFF + i * STRIDE] = t[i];
        });
    }
}

}  // namespace detail
}  // namespace smfft
