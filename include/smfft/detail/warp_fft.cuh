// smfft/detail/warp_fft.cuh -- in-place FFT stages on 4 points per thread with WARP-SHUFFLE exchanges.
//
// The engine behind the reference-contract device API (include/smfft/compat.cuh) where that contract -- four points
// per thread, fft_length/4 threads, tile in natural order in shared memory -- lets a whole (sub-)transform live inside
// one warp: north_star subsystem (2), "register-resident radix-4 stages with warp-shuffle exchanges for intra-warp
// strides and shared memory only across warps".  It replaces, for those cases, the reference's schedule of one radix-2
// stage per shuffle round (8 SHFL per stage and thread, CT/FFT-GPU-32bit.cu:363-411) and its reorder_* permutations
// (16-24 SHFL + 4-8 LDS/STS + 3-5 BAR, CT:54-329): a radix-4 stage here costs ONE 4x4 transposition between the
// register index and two lane bits (6 SHFL), and the bit reversal is the addressing of the one store that follows.
//
// Model.  A tile position has bits Q0..Q(n-1).  Every bit lives in a SLOT: R0, R1 (the two bits of the register index
// m of v[m]), L0..L4 (lane bits) or W0..W4 (warp-index bits).  A PLAN is an initial slot map plus a list of steps:
//   X4<a,b>  transpose the register index with lane bits (a,b): slot R0 <-> La, R1 <-> Lb          (6 SHFL)
//   X2<r,l>  swap register bit r with lane bit l                                                      (4 SHFL)
//   S4       radix-4 stage over the position bits held by (R1,R0)  (must be consecutive bits hi = lo + 1)
//   S2<r>    radix-2 stage over the position bit held by register bit r
// Two in-place algorithms run on such plans (formulas checked against FP64 in tests/test_warp_fft_model.py):
//   DIF (fft_reorder = 1): natural-order input, stages from the TOP position bit down; a stage over bits (hi,lo) takes
//        input digit a = 2 P_hi + P_lo, multiplies output k by W_{2^(hi+1)}^{k * low} (low = the bits below lo) and
//        stores k bit-swapped, (P_hi,P_lo) = (k0,k1); at the end position p holds X[brev(p)], so the ONE store that
//        follows writes to address brev(p): the reorder_* pass of the reference is that store's addressing;
//   DIT (fft_reorder = 0): computes DFT(x o brev) in natural order, stages from the BOTTOM bit up; a stage multiplies
//        input a = bitswap(position digit) by W_{2^(hi+1)}^{a * klow} (klow = the bits below lo, already outputs) and
//        stores output k naturally.
// Twiddles: one MUFU sincos per stage and thread (the reference's source, CT:18-28; this API has no table pointer),
// higher powers by complex multiplication.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "layout.cuh"
#include "radix.cuh"
#include "twiddle.cuh"
#include "warp_xchg.cuh"

namespace smfft {
namespace detail {
namespace wf {

constexpr int kSlots = 12;  // R0 R1 | L0..L4 | W0..W4
// packed f32x2 add / subtract on (re, im) in the butterflies: these engines are issue-bound (about 60 instructions per
// point, profiles/r02_compat_*), and FADD2 halves the additions; results are bit-identical
#ifndef SMFFT_WF_PACK
#define SMFFT_WF_PACK 1
#endif
constexpr int kPack = SMFFT_WF_PACK;
enum { SL_R0 = 0, SL_R1 = 1, SL_L0 = 2, SL_W0 = 7 };
enum { OP_X4 = 0, OP_X2 = 1, OP_S4 = 2, OP_S2 = 3 };

struct SlotMap {
    int q[kSlots];  // q[slot] = tile-position bit held by the slot, -1 = slot unused
};
struct Op {
    int kind, a, b;
};

SMFFT_CX SlotMap apply_op(SlotMap m, Op o)
{
    if (o.kind == OP_X4) {
        const int t0 = m.q[SL_R0], t1 = m.q[SL_R1];
        m.q[SL_R0] = m.q[SL_L0 + o.a];
        m.q[SL_R1] = m.q[SL_L0 + o.b];
        m.q[SL_L0 + o.a] = t0;
        m.q[SL_L0 + o.b] = t1;
    } else if (o.kind == OP_X2) {
        const int t = m.q[o.a];
        m.q[o.a] = m.q[SL_L0 + o.b];
        m.q[SL_L0 + o.b] = t;
    }
    return m;
}

// map after the first `n` steps of plan P
template <class P>
SMFFT_CX SlotMap map_after(int n)
{
    SlotMap m = P::init();
    for (int i = 0; i < n; i++) m = apply_op(m, P::op(i));
    return m;
}

SMFFT_CX int bitswap2(int d) { return ((d & 1) << 1) | (d >> 1); }

// where tile-position bit q goes in a shared / global address:  natural, or bit-reversed inside the low E bits (DIF result)
template <int E, bool BREV>
SMFFT_CX int dst_bit(int q)
{
    return (BREV && q < E) ? (E - 1 - q) : q;
}

struct Lane {
    int lane, warp;
};
SMFFT_DEV Lane whoami()
{
    const int t = plat::tid();
    return Lane{t & 31, t >> 5};
}

// value of the position bits [LO, HI) that live in lane / warp slots, each moved to dst_bit()
template <class MAP_HOLDER, int LO, int HI, int E, bool BREV>
SMFFT_DEV int thread_bits(Lane w)
{
    constexpr SlotMap M = MAP_HOLDER::value();
    int r = 0;
    static_for<kSlots - 2>([&](auto SI) {
        constexpr int s = 2 + decltype(SI)::value;
        constexpr int q = M.q[s];
        if constexpr (q >= LO && q < HI) {
            constexpr int d = dst_bit<E, BREV>(q);
            const int src = s < SL_W0 ? w.lane : w.warp;
            constexpr int sb = s < SL_W0 ? s - SL_L0 : s - SL_W0;
            if constexpr (d >= sb)
                r |= (src & (1 << sb)) << (d - sb);
            else
                r |= (src & (1 << sb)) >> (sb - d);
        }
    });
    return r;
}
// the same for the bits that live in the register index m (compile time)
template <int LO, int HI, int E, bool BREV>
SMFFT_CX int reg_bits(SlotMap M, int m)
{
    int r = 0;
    for (int s = 0; s < 2; s++) {
        const int q = M.q[s];
        if (q >= LO && q < HI) r |= ((m >> s) & 1) << dst_bit<E, BREV>(q);
    }
    return r;
}

template <class P, int N>
struct MapAt {
    static SMFFT_CX SlotMap value() { return map_after<P>(N); }
};

// ---- stages ------------------------------------------------------------------------------------------------------

// W_WN^{m} with W = exp(-+ 2 pi i / WN); constants for register-held bits come from mul_wconst (moduli up to 64)
template <int DIR, int WN>
SMFFT_DEV float2 tw_base(int m)
{
    return tw_mufu<DIR, WN>(m);
}

// radix-4 stage over the bits held by (R1, R0) under map M (= the map BEFORE the stage; stages do not move bits)
template <class P, int STEP, int DIR>
SMFFT_DEV void stage4(float2 (&v)[4], Lane w)
{
    constexpr SlotMap M = map_after<P>(STEP);
    constexpr int lo = M.q[SL_R0], hi = M.q[SL_R1];
    static_assert(hi == lo + 1 && lo >= 0 && hi < P::E, "radix-4 stage: (R1,R0) must hold consecutive transform bits");
    constexpr int WN = 1 << (hi + 1);
    if constexpr (P::DIT) {
        if constexpr (lo > 0) {
            const int klow = thread_bits<MapAt<P, STEP>, 0, lo, P::E, false>(w);
            const float2 w1 = tw_base<DIR, WN>(klow), w2 = csqr(w1), w3 = cmul(w1, w2);
            // register m holds position digit m, i.e. input a = bitswap(m)
            v[2] = cmul(v[2], w1);
            v[1] = cmul(v[1], w2);
            v[3] = cmul(v[3], w3);
        }
        const float2 t = v[1];  // order the inputs by a
        v[1] = v[2];
        v[2] = t;
        dft_regs<kPack, DIR, 4, 0, 1, 4>(v);
    } else {
        dft_regs<kPack, DIR, 4, 0, 1, 4>(v);
        if constexpr (lo > 0) {
            const int low = thread_bits<MapAt<P, STEP>, 0, lo, P::E, false>(w);
            const float2 w1 = tw_base<DIR, WN>(low), w2 = csqr(w1), w3 = cmul(w1, w2);
            v[1] = cmul(v[1], w1);
            v[2] = cmul(v[2], w2);
            v[3] = cmul(v[3], w3);
        }
        const float2 t = v[1];  // output k is stored bit-swapped
        v[1] = v[2];
        v[2] = t;
    }
}

// radix-2 stage over the bit held by register bit RB; the other register bit may hold a lower transform bit (then it
// adds a compile-time factor to the twiddle), a higher one, or a transform-index bit (two independent butterflies)
template <class P, int STEP, int DIR, int RB>
SMFFT_DEV void stage2(float2 (&v)[4], Lane w)
{
    constexpr SlotMap M = map_after<P>(STEP);
    constexpr int h = M.q[RB], g = M.q[1 - RB];
    static_assert(h >= 0 && h < P::E, "radix-2 stage: the register bit must hold a transform bit");
    constexpr int WN = 1 << (h + 1);
    constexpr bool other_low = g >= 0 && g < h;
    static_assert(!other_low || h + 1 - g <= 6, "constant twiddle factors come from the W_64 table");
    float2 wl = make_float2(1.0f, 0.0f);
    constexpr bool has_tw = h > 0;
    if constexpr (has_tw) {
        const int low = thread_bits<MapAt<P, STEP>, 0, h, P::E, false>(w);
        wl = tw_base<DIR, WN>(low);
    }
    static_for<2>([&](auto OI) {
        constexpr int o = decltype(OI)::value;
        constexpr int m0 = o << (1 - RB), m1 = m0 | (1 << RB);
        constexpr int DEN = other_low ? (1 << (h + 1 - g)) : 1;  // W_WN^{o 2^g} = W_DEN^{o}
        if constexpr (P::DIT) {
            float2 b = v[m1];
            if constexpr (has_tw) b = cmul(b, wl);
            if constexpr (other_low && o == 1) b = mul_wconst<DIR, 1, DEN>(b);
            const float2 a = v[m0];
            v[m0] = cadd_p<kPack>(a, b);
            v[m1] = csub_p<kPack>(a, b);
        } else {
            const float2 a = v[m0], b = v[m1];
            v[m0] = cadd_p<kPack>(a, b);
            float2 d = csub_p<kPack>(a, b);
            if constexpr (has_tw) d = cmul(d, wl);
            if constexpr (other_low && o == 1) d = mul_wconst<DIR, 1, DEN>(d);
            v[m1] = d;
        }
    });
}

template <class P, int DIR, int FIRST = 0, int LAST = P::NOPS>
SMFFT_DEV void run_plan(float2 (&v)[4], Lane w)
{
    static_for<LAST - FIRST>([&](auto SI) {
        constexpr int i = FIRST + decltype(SI)::value;
        constexpr Op o = P::op(i);
        if constexpr (o.kind == OP_X4)
            xchg4<o.a, o.b>(v, w.lane);
        else if constexpr (o.kind == OP_X2)
            xchg2<o.a, o.b>(v, w.lane);
        else if constexpr (o.kind == OP_S4)
            stage4<P, i, DIR>(v, w);
        else
            stage2<P, i, DIR, o.a>(v, w);
    });
}

// ---- tile <-> registers by slot map --------------------------------------------------------------------------------

// address (in float2 units, before the layout) of register m of this thread: position bits -> address bits
template <class P, int STEP, bool BREV, class LY, class F>
SMFFT_DEV void for_each_reg(Lane w, F&& f)
{
    constexpr SlotMap M = map_after<P>(STEP);
    const int base = thread_bits<MapAt<P, STEP>, 0, 32, P::E, BREV>(w);
    static_for<4>([&](auto MI) {
        constexpr int m = decltype(MI)::value;
        constexpr int off = reg_bits<0, 32, P::E, BREV>(M, m);
        f(MI, LY::phys(base | off));
    });
}

template <class P, int STEP, class LY = LayoutLinear, class PTR>
SMFFT_DEV void load_by_map(float2 (&v)[4], PTR s, Lane w)
{
    for_each_reg<P, STEP, false, LY>(w, [&](auto MI, int a) { v[decltype(MI)::value] = plat::lds64(s + a); });
}
// BREV: the DIF result (position p holds X[brev(p)]) goes to its natural address
template <class P, int STEP, bool BREV, class LY = LayoutLinear, class PTR>
SMFFT_DEV void store_by_map(const float2 (&v)[4], PTR s, Lane w)
{
    for_each_reg<P, STEP, BREV, LY>(w, [&](auto MI, int a) { plat::sts64(s + a, v[decltype(MI)::value]); });
}

// First read of the DIT plans: thread t owns the four CONSECUTIVE positions 4t .. 4t+3 (R0 = Q0, R1 = Q1, lanes and
// warps the bits above).  Read as 4 x 64 bit with the register index rotated by (lane >> 2) & 3, so that the sixteen
// lanes of a half-warp touch sixteen different bank pairs (a plain run of four would collide 4-way), then rotated back
// in registers: 4 LDS.64 + 16 SEL, no alignment demand on the tile.
SMFFT_DEV void load_rows_rotated(float2 (&v)[4], const float2* s, int t, int lane)
{
    if ((plat::smem_addr(s) & 15) == 0) {
        // 16-byte aligned tile: two 128-bit reads; lanes 4..7 of every eight fetch their halves in the other order so
        // that a quarter-warp touches eight different 16-byte bank groups, and swap them back (8 SEL)
        const bool hs = (lane >> 2) & 1;
        const float4 a = plat::lds128(s + 4 * t + (hs ? 2 : 0)), b = plat::lds128(s + 4 * t + (hs ? 0 : 2));
        v[0] = hs ? make_float2(b.x, b.y) : make_float2(a.x, a.y);
        v[1] = hs ? make_float2(b.z, b.w) : make_float2(a.z, a.w);
        v[2] = hs ? make_float2(a.x, a.y) : make_float2(b.x, b.y);
        v[3] = hs ? make_float2(a.z, a.w) : make_float2(b.z, b.w);
        return;
    }
    const int c = (lane >> 2) & 3;
    float2 a[4];
    static_for<4>([&](auto MI) {
        constexpr int m = decltype(MI)::value;
        a[m] = plat::lds64(s + 4 * t + ((m + c) & 3));  // a[m] = x[(m + c) & 3]
    });
    // x[j] = a[(j - c) & 3]: undo a rotation by 1, then by 2
    const bool c0 = c & 1, c1 = c & 2;
    float2 b[4];
    b[0] = c0 ? a[3] : a[0];
    b[1] = c0 ? a[0] : a[1];
    b[2] = c0 ? a[1] : a[2];
    b[3] = c0 ? a[2] : a[3];
    v[0] = c1 ? b[2] : b[0];
    v[1] = c1 ? b[3] : b[1];
    v[2] = c1 ? b[0] : b[2];
    v[3] = c1 ? b[1] : b[3];
}

// ---- plans ---------------------------------------------------------------------------------------------------------
// Maps are written as {R0, R1, L0, L1, L2, L3, L4, W0, W1, W2, W3, W4}.

// one warp, 128-point tile = 128 >> E transforms of 2^E points, natural order in and out (fft_reorder = 1), DIF
template <int E>
struct PlanDif;
template <>
struct PlanDif<7> {  // column ownership: conflict-free read; the result lands with k0..k3 in lane bits: conflict-free write
    static constexpr int E = 7, NOPS = 7;
    static constexpr bool DIT = false;
    static SMFFT_CX SlotMap init() { return SlotMap{{5, 6, 0, 1, 2, 3, 4, -1, -1, -1, -1, -1}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 3, 4}, {OP_S4, 0, 0}, {OP_X4, 1, 2}, {OP_S4, 0, 0}, {OP_X2, 0, 0}, {OP_S2, 0, 0}};
        return t[i];
    }
};
template <>
struct PlanDif<6> {  // two transforms: lanes 0-15 / 16-31
    static constexpr int E = 6, NOPS = 5;
    static constexpr bool DIT = false;
    static SMFFT_CX SlotMap init() { return SlotMap{{4, 5, 0, 1, 2, 3, 6, -1, -1, -1, -1, -1}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 2, 3}, {OP_S4, 0, 0}, {OP_X4, 0, 1}, {OP_S4, 0, 0}};
        return t[i];
    }
};
template <>
struct PlanDif<5> {  // four transforms: register bit 1 and lane bit 4 select them; [2,4,4]
    static constexpr int E = 5, NOPS = 5;
    static constexpr bool DIT = false;
    static SMFFT_CX SlotMap init() { return SlotMap{{4, 6, 0, 1, 2, 3, 5, -1, -1, -1, -1, -1}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S2, 0, 0}, {OP_X4, 2, 3}, {OP_S4, 0, 0}, {OP_X4, 0, 1}, {OP_S4, 0, 0}};
        return t[i];
    }
};

// one warp, 128 consecutive positions, DFT of the bit-reversed input in natural order (fft_reorder = 0), DIT.
// First read by rows (load_rows_rotated): R0 = Q0, R1 = Q1.  For E = 7 this is also PHASE 1 of every larger
// no-reorder transform: its first seven stages are exactly the 128-point transforms of the 128-position blocks.
template <int E>
struct PlanDit;
template <>
struct PlanDit<7> {
    static constexpr int E = 7, NOPS = 7;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 0, 1}, {OP_S4, 0, 0}, {OP_X4, 2, 3}, {OP_S4, 0, 0}, {OP_X2, 0, 4}, {OP_S2, 0, 0}};
        return t[i];
    }
};
template <>
struct PlanDit<6> {
    static constexpr int E = 6, NOPS = 5;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 0, 1}, {OP_S4, 0, 0}, {OP_X4, 2, 3}, {OP_S4, 0, 0}};
        return t[i];
    }
};
template <>
struct PlanDit<5> {  // [4,4,2]; the last bit comes in by a one-bit swap (4 SHFL + 12 SEL).  A second 4x4 transposition (6 SHFL + 32 SEL)
                     // would also bring a transform bit into the registers and make the final store conflict-free, but these one-warp
                     // engines are issue-bound (ncu: 74 % issue): 2 extra store wavefronts are cheaper than 22 more instructions
    static constexpr int E = 5, NOPS = 5;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 0, 1}, {OP_S4, 0, 0}, {OP_X2, 0, 2}, {OP_S2, 0, 0}};
        return t[i];
    }
};

// PHASE 2 of the larger no-reorder transforms (E = 8..12): the bits 7..E-1, read back from shared memory with the low
// address bits in lane slots (whole 128-byte rows per half-warp while E <= 10; 64- and 32-byte pieces for 2048 and 4096
// points, where the exchange layout un-conflicts the read and only the final, linear store pays 2 and 4 wavefronts)
template <int E>
struct PlanDit2;
template <>
struct PlanDit2<8> {
    static constexpr int E = 8, NOPS = 1;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{7, 6, 0, 1, 2, 3, 5, 4, -1, -1, -1, -1}}; }
    static SMFFT_CX Op op(int) { return Op{OP_S2, 0, 0}; }
    using XLayout = LayoutLinear;
};
template <>
struct PlanDit2<9> {
    static constexpr int E = 9, NOPS = 1;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{7, 8, 0, 1, 2, 3, 4, 5, 6, -1, -1, -1}}; }
    static SMFFT_CX Op op(int) { return Op{OP_S4, 0, 0}; }
    using XLayout = LayoutLinear;
};
template <>
struct PlanDit2<10> {
    static constexpr int E = 10, NOPS = 3;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{7, 8, 0, 1, 2, 3, 9, 4, 5, 6, -1, -1}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X2, 0, 4}, {OP_S2, 0, 0}};
        return t[i];
    }
    using XLayout = LayoutLinear;
};
// exchange layouts of the two largest sizes: the address bits the half-warp does NOT cover are folded into the bank bits
struct LayoutX11 {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 9) & 1) << 3); }
};
struct LayoutX12 {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 9) & 3) << 2); }
};
template <>
struct PlanDit2<11> {
    static constexpr int E = 11, NOPS = 3;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{7, 8, 0, 1, 2, 9, 10, 3, 4, 5, 6, -1}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 3, 4}, {OP_S4, 0, 0}};
        return t[i];
    }
    using XLayout = LayoutX11;
};
template <>
struct PlanDit2<12> {
    static constexpr int E = 12, NOPS = 5;
    static constexpr bool DIT = true;
    static SMFFT_CX SlotMap init() { return SlotMap{{7, 8, 0, 1, 9, 10, 11, 2, 3, 4, 5, 6}}; }
    static SMFFT_CX Op op(int i)
    {
        constexpr Op t[NOPS] = {{OP_S4, 0, 0}, {OP_X4, 2, 3}, {OP_S4, 0, 0}, {OP_X2, 0, 4}, {OP_S2, 0, 0}};
        return t[i];
    }
    using XLayout = LayoutX12;
};

// ---- transforms on a tile in shared memory (linear, natural order in and out, in place) --------------------------

// 128-point tile held by ONE warp: 128 >> E transforms of 2^E points (E = 5, 6, 7).  No block barrier inside.
template <int E, int DIR, int REORDER>
SMFFT_DEV void warp_tile_fft(float2* s)
{
    const Lane w = Lane{plat::tid() & 31, 0};
    float2 v[4];
    if constexpr (REORDER) {
        using P = PlanDif<E>;
        load_by_map<P, 0>(v, s, w);
        run_plan<P, DIR>(v, w);
        plat::sync_warp();  // every lane has read its inputs before results overwrite them
        store_by_map<P, P::NOPS, true>(v, s, w);
    } else {
        using P = PlanDit<E>;
        load_rows_rotated(v, s, w.lane, w.lane);
        run_plan<P, DIR>(v, w);
        plat::sync_warp();
        store_by_map<P, P::NOPS, false>(v, s, w);
    }
}

// 2^E points (E = 8..12), fft_reorder = 0, 2^E / 4 threads: phase 1 = PlanDit<7> per warp on its 128 consecutive
// positions, one exchange through the tile (in place), phase 2 = PlanDit2<E>.  ONE block barrier inside; every warp
// reads and writes only its own positions in each phase (a second barrier only where the exchange is swizzled, 2048 and 4096 points).
template <int E, int DIR>
SMFFT_DEV void block_fft_noreorder(float2* s)
{
    const Lane w = whoami();
    float2 v[4];
    using P1 = PlanDit<7>;
    using P2 = PlanDit2<E>;
    using XL = typename P2::XLayout;
    load_rows_rotated(v, s, plat::tid(), w.lane);
    run_plan<P1, DIR>(v, w);
    plat::sync_warp();
    store_by_map<P1, P1::NOPS, false, XL>(v, s, w);
    plat::sync_block();
    load_by_map<P2, 0, XL>(v, s, w);
    // a swizzled exchange keeps position x at address x ^ key, which is a slot ANOTHER warp writes its (linear) result to:
    // every warp must have read before any result lands.  With the linear exchange each warp reads and writes its own slots.
    if constexpr (!std::is_same<XL, LayoutLinear>::value) plat::sync_block();
    run_plan<P2, DIR>(v, w);
    plat::sync_warp();
    store_by_map<P2, P2::NOPS, false>(v, s, w);
}

// ---- the same transforms with the tile in GLOBAL memory (the external wrapper kernels: no staging copy) --------------

// thread t's four consecutive points 4t .. 4t+3 of the tile at g: two 16-byte loads when the tile is 16-byte aligned
SMFFT_DEV void load_rows_global(float2 (&v)[4], const float2* __restrict__ g, int t)
{
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const float4 a = plat::ldg128_stream(g + 4 * t), b = plat::ldg128_stream(g + 4 * t + 2);
        v[0] = make_float2(a.x, a.y);
        v[1] = make_float2(a.z, a.w);
        v[2] = make_float2(b.x, b.y);
        v[3] = make_float2(b.z, b.w);
    } else {
        static_for<4>([&](auto MI) { v[decltype(MI)::value] = plat::ldg64_stream(g + 4 * t + decltype(MI)::value); });
    }
}

// one warp, 128-point tile: global -> registers -> shuffles -> global; no shared memory, no barrier
template <int E, int DIR, int REORDER>
SMFFT_DEV void warp_tile_fft_global(const float2* __restrict__ gin, float2* __restrict__ gout)
{
    const Lane w = Lane{plat::tid() & 31, 0};
    float2 v[4];
    if constexpr (REORDER) {
        using P = PlanDif<E>;
        for_each_reg<P, 0, false, LayoutLinear>(w, [&](auto MI, int a) { v[decltype(MI)::value] = plat::ldg64_stream(gin + a); });
        run_plan<P, DIR>(v, w);
        for_each_reg<P, P::NOPS, true, LayoutLinear>(w, [&](auto MI, int a) { plat::stg64_stream(gout + a, v[decltype(MI)::value]); });
    } else {
        using P = PlanDit<E>;
        load_rows_global(v, gin, w.lane);
        run_plan<P, DIR>(v, w);
        for_each_reg<P, P::NOPS, false, LayoutLinear>(w, [&](auto MI, int a) { plat::stg64_stream(gout + a, v[decltype(MI)::value]); });
    }
}

// 2^E points (E = 8..12), fft_reorder = 0: rows straight from global memory, shared memory only for the one exchange
template <int E, int DIR>
SMFFT_DEV void block_fft_noreorder_global(float2* s, const float2* __restrict__ gin, float2* __restrict__ gout)
{
    const Lane w = whoami();
    float2 v[4];
    using P1 = PlanDit<7>;
    using P2 = PlanDit2<E>;
    using XL = typename P2::XLayout;
    load_rows_global(v, gin, plat::tid());
    run_plan<P1, DIR>(v, w);
    store_by_map<P1, P1::NOPS, false, XL>(v, s, w);
    plat::sync_block();
    load_by_map<P2, 0, XL>(v, s, w);
    run_plan<P2, DIR>(v, w);
    for_each_reg<P2, P2::NOPS, false, LayoutLinear>(w, [&](auto MI, int a) { plat::stg64_stream(gout + a, v[decltype(MI)::value]); });
}

}  // namespace wf
}  // namespace detail
}  // namespace smfft
