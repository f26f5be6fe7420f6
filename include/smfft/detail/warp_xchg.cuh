// smfft/detail/warp_xchg.cuh -- register <-> lane transpositions by warp shuffle (shared by warp_fft.cuh and block_fft.cuh).
//
// The reference exchanges one radix-2 stage at a time (shfl / shfl_xor helpers, CT/FFT-GPU-32bit.cu:30-44, 8 SHFL per
// stage and thread, CT:363-411).  Here an exchange moves TWO index bits at once: a 4x4 transposition between four
// registers and the four lanes that differ in two lane bits costs 6 SHFL (three of the four registers travel), a 2x2
// swap of one register bit with one lane bit 4 SHFL per four registers.
#pragma once
#include "complex.cuh"

namespace smfft {
namespace detail {
namespace wf {

SMFFT_DEV float2 shfl_xor2(float2 v, int mask) { return make_float2(plat::shfl_xor(v.x, mask), plat::shfl_xor(v.y, mask)); }
SMFFT_DEV void cswap(bool c, float2& a, float2& b)
{
    const float2 ta = c ? b : a, tb = c ? a : b;
    a = ta;
    b = tb;
}

// 4x4 transposition between the register index and the lane digit d = bit A | bit B << 1:
// afterwards v[m] on the lane with digit c is what v[c] was on the lane with digit m.
// Round r = 1..3 trades register (d ^ r) with the lane whose digit is d ^ r; holding the registers XOR-permuted by d
// (two conditional-swap levels before and after) makes every register index in the shuffles static.
template <int A, int B>
SMFFT_DEV void xchg4(float2 (&v)[4], int lane)
{
    const bool d0 = (lane >> A) & 1, d1 = (lane >> B) & 1;
    cswap(d0, v[0], v[1]);
    cswap(d0, v[2], v[3]);
    cswap(d1, v[0], v[2]);
    cswap(d1, v[1], v[3]);
    v[1] = shfl_xor2(v[1], 1 << A);
    v[2] = shfl_xor2(v[2], 1 << B);
    v[3] = shfl_xor2(v[3], (1 << A) | (1 << B));
    cswap(d0, v[0], v[1]);
    cswap(d0, v[2], v[3]);
    cswap(d1, v[0], v[2]);
    cswap(d1, v[1], v[3]);
}

// swap register bit RB with lane bit LB: the lane keeps the two registers whose bit RB equals its lane bit and trades the others
template <int RB, int LB>
SMFFT_DEV void xchg2(float2 (&v)[4], int lane)
{
    const bool b = (lane >> LB) & 1;
    static_for<2>([&](auto OI) {
        constexpr int o = decltype(OI)::value;
        constexpr int m0 = o << (1 - RB), m1 = m0 | (1 << RB);
        const float2 send = b ? v[m0] : v[m1];
        const float2 got = shfl_xor2(send, 1 << LB);
        if (b) v[m0] = got; else v[m1] = got;
    });
}

}  // namespace wf
}  // namespace detail
}  // namespace smfft
