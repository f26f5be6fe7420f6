// smfft/detail/block_fft.cuh -- the native block FFT: F independent N-point C2C FFTs on one tile
// of shared memory, R points per thread, register-resident radix-R passes.
//
// Replaces, on the hot path, the reference's per-stage schedules:
//   do_SMFFT_CT_DIT<P>              CT/FFT-GPU-32bit.cu:334-532  (+ reorder_* 54-329)
//   do_FFT_Stockham_mk6<P>          ST/FFT-GPU-32bit-Stockham.cu:97-240
//   do_FFT_Stockham_C2C<P,Dir>      RC/FFT-GPU-32bit-Stockham.cu:106-266
//   do_FFT_Stockham_R2C_C2R<P,Dir>  RC/FFT-GPU-32bit-Stockham.cu:269-344
//
// Algorithm (Stockham autosort, mixed radix r_0 .. r_{P-1}, all powers of two, r_p <= R):
//   thread t of an FFT (T = N/R threads) always READS the R points  x = t + m*T, m = 0..R-1;
//   a pass of radix r treats them as U = R/r butterflies: butterfly u = "virtual thread"
//   j = t + u*T, its q-th input is register m = u + q*U;
//   twiddle  v[u+qU] *= W_{Ns*r}^{(j mod Ns) q}      (Ns = r_0*..*r_{p-1}, nothing for Ns = 1)
//   DFT_r in registers, natural order out
//   WRITES   x = (j div Ns)*Ns*r + (j mod Ns) + q*Ns (the autosort step)
//   The last pass writes x = t + m*T again, i.e. exactly the slots the thread read, so the
//   transform is in place on the tile and the final pass needs no barrier.
// Output is the natural-order un-normalised DFT (CT reorder=1, Stockham).
//
// fft_reorder = 0 (output = DFT of the bit-reversed input): the only change
// is the first read.  brev_e(t + m*T) = brev_a(t)*R + brev_b(m), so a thread reads ONE contiguous
// run of R points (128-bit loads) in bit-reversed register order -- the permutation that costs the
// reference 16-24 SHFL + 4-8 LDS/STS + 3-5 BAR per thread is register renaming
// here, fused into the load.  To keep both that read and the first exchange write conflict-free
// under SW128, thread t acts as virtual thread j(t) = t ^ ((brev3(t&7) << (a-3)) & ~7): within any
// 8 consecutive lanes both j & 7 and brev_a(j) & 7 take 8 distinct values.
#pragma once
#include <type_traits>

#include "layout.cuh"
#include "radix.cuh"
#include "twiddle.cuh"
#include "warp_xchg.cuh"

namespace smfft {
namespace detail {

// Layout_  : layout of the tile on entry and exit (SW128 for the TMA kernels, linear for the
//            reference-compatible device API);  XLayout_: layout used by the exchanges between passes.
// VEC128_  : allow 16-byte shared accesses (needs a 16-byte aligned tile).
// SKEW_    : de-conflict the natural accesses of small transforms (T < 16) with skewed rows + selects.
// DUAL_    : bit flags.  2 = packed f32x2 add / subtract on (re, im) in the butterflies (same threads, fewer issue slots);
//            4 = reversed pass plan (small radix first; lets the C2R pass own its pairs, see MirrorC2R);
//            1 = one thread carries the same points of TWO transforms of the tile in the packed f32x2 lanes
//            (block_fft_dual.cuh): half the threads, half the floating-point and exchange instructions per point;
//            8 = two-pass plans of transforms held by 2 or 4 LANES (32 / 64 points at R = 16) exchange through warp
//            shuffles instead of the tile (warp_xchg.cuh): no barrier, a third of the shared-memory traffic.
template <int E_, int B_, int F_, int DIR_, int REORDER_, int TW_, class Layout_ = LayoutSW128, class XLayout_ = Layout_,
          bool VEC128_ = true, bool SKEW_ = true, int DUAL_ = 0>
struct BlockCfg {
    static constexpr int E = E_;        // log2 N
    static constexpr int N = 1 << E_;   // FFT length
    static constexpr int B = B_;        // log2 R
    static constexpr int R = 1 << B_;   // points per thread
    static constexpr int A = E_ - B_;   // log2 T
    static constexpr int T = 1 << A;    // threads per FFT
    static constexpr int F = F_;        // FFTs per tile (power of two)
    static constexpr int L = F_ * N;    // points per tile
    static constexpr int DUAL = DUAL_ & 1;
    static constexpr int PACK = (DUAL_ >> 1) & 1;  // one transform per thread, packed (re, im) add / subtract in the butterflies
    static constexpr int REV = (DUAL_ >> 2) & 1;   // pass plan with the small radix FIRST: [2^(E mod B), R, .., R]
    static constexpr int XSHFL = (DUAL_ >> 3) & 1;  // the one exchange of a two-pass plan by warp shuffles (T = 2 or 4 lanes per transform)
    static constexpr int THREADS = (F_ * T) >> DUAL;
    static_assert(DUAL_ >= 0 && DUAL_ < 16 && (!REV || ((E_ % B_ == 3 || E_ % B_ == 2) && VEC128_)), "reversed plans: first radix 8 or 4");
    static_assert(((DUAL_ >> 3) & 1) == 0 || ((DUAL_ & 5) == 0 && B_ == 4 && (E_ == 5 || E_ == 6)), "shuffle exchange: R = 16, 32 or 64 points, plan [16, T]");  // 64 points: built and measured, not used (tuning.hpp)
    static_assert((DUAL_ & 1) == 0 || (DUAL_ == 1 && F_ % 2 == 0 && E_ - B_ >= 4 && E_ >= 7 && B_ >= 4 && (E_ + B_ - 1) / B_ >= 2 &&
                                 std::is_same<Layout_, LayoutSW128>::value && VEC128_),
                  "dual-lane transforms: an even number of transforms per tile, T >= 16, N >= 128, R >= 16, SW128 tile");
    static constexpr int DIR = DIR_;          // 0 forward (exp -), 1 inverse (exp +): FFT_Params::fft_direction
    static constexpr int REORDER = REORDER_;  // FFT_Params::fft_reorder
    static constexpr int TW = TW_;
    using Layout = Layout_;
    using XLayout = XLayout_;
    static constexpr bool VEC128 = VEC128_;
    static constexpr bool SKEW_SMALL = SKEW_;
    static constexpr bool SAME_LAYOUT = std::is_same<Layout_, XLayout_>::value;
    static_assert(E_ > B_, "need at least two threads per FFT");
    static_assert(B_ >= 1 && B_ <= 5, "1..32 points per thread");
    // pass plan, large radix first: [R, R, .., R, 2^(E mod B)]  (REV: [2^(E mod B), R, .., R])
    static constexpr int P = (E + B - 1) / B;
    static SMFFT_CX int radix_log2(int p) { return (REV && E % B != 0) ? (p == 0 ? E % B : B) : (p < E / B ? B : E % B); }
    static SMFFT_CX int ns_log2(int p)
    {
        int s = 0;
        for (int i = 0; i < p; i++) s += radix_log2(i);
        return s;
    }
    // compact twiddle table (TW_LUT): pass p > 0 owns Ns_p entries W_{Ns_p r_p}^k at tw_offset(p);
    // the R2C/C2R pass owns T entries W_{2N}^t after them (W_{2N}^{t+mT} = W_{2N}^t * W_{2R}^m)
    static SMFFT_CX int tw_offset(int p)
    {
        int o = 0;
        for (int i = 1; i < p; i++) o += 1 << ns_log2(i);
        return o;
    }
    static constexpr int TW_C2C_ENTRIES = tw_offset((E_ + B_ - 1) / B_);
    static constexpr int TW_R2C_ENTRIES = 1 << (E_ - B_);
};

// Fill the compact table from the global FP64-rounded table gtw[j] = exp(-2 pi i j / kTwiddleTableSize).
// WITH_R2C: also the pair-pass twiddles (direction R2C_INVERSE).  Caller synchronises afterwards.
template <class C, bool WITH_R2C, int R2C_INVERSE>
SMFFT_DEV void fill_twiddle_table(float2* stw, const float2* __restrict__ gtw, int tid, int nthreads)
{
    if constexpr (C::TW == TW_LUT) {
        static_for<C::P>([&](auto PI) {
            constexpr int p = decltype(PI)::value;
            if constexpr (p >= 1) {
                constexpr int NS = 1 << C::ns_log2(p), WN = NS << C::radix_log2(p);
                for (int k = tid; k < NS; k += nthreads) {
                    float2 w = plat::ldg_ro(gtw + k * (kTwiddleTableSize / WN));
                    if (C::DIR) w.y = -w.y;
                    stw[C::tw_offset(p) + k] = w;
                }
            }
        });
        if constexpr (WITH_R2C) {
            for (int k = tid; k < C::TW_R2C_ENTRIES; k += nthreads) {
                float2 w = plat::ldg_ro(gtw + k * (kTwiddleTableSize / (2 * C::N)));
                if (R2C_INVERSE) w.y = -w.y;
                stw[C::TW_C2C_ENTRIES + k] = make_float2(0.5f * w.x, 0.5f * w.y);  // W_{2N}^t / 2, t < T (real_pass_regs)
            }
        }
    }
}

// The same fill in two phases for short-lived CTAs (register-direct kernels, one tile per CTA): the global reads are
// ISSUED early (right after the CTA's own input loads) into registers, and only stored to the shared table later, behind
// the first pass's arithmetic -- a table read that precedes the input loads costs such a CTA a whole memory round trip.
// C2C tables only; K = entries per thread.
template <class C>
struct TwiddlePrefetch {
    static constexpr int K = C::TW == TW_LUT ? (C::TW_C2C_ENTRIES + C::THREADS - 1) / C::THREADS : 0;
    static constexpr bool OK = K >= 0 && K <= 4;
    float2 w[K > 0 ? K : 1];
    // entry i of the compact table = pass p (>= 1), index k inside it
    static SMFFT_DEV void locate(int i, int& stride, bool& valid)
    {
        valid = false;
        stride = 0;
        static_for<C::P>([&](auto PI) {
            constexpr int p = decltype(PI)::value;
            if constexpr (p >= 1) {
                constexpr int NS = 1 << C::ns_log2(p), WN = NS << C::radix_log2(p), OFF = C::tw_offset(p);
                if (i >= OFF && i < OFF + NS) {
                    valid = true;
                    stride = (i - OFF) * (kTwiddleTableSize / WN);
                }
            }
        });
    }
    SMFFT_DEV void load(const float2* __restrict__ gtw, int tid)
    {
        static_for<(K > 0 ? K : 0)>([&](auto KI) {
            constexpr int j = decltype(KI)::value;
            int stride;
            bool valid;
            locate(tid + j * C::THREADS, stride, valid);
            w[j] = valid ? plat::ldg_ro(gtw + stride) : make_float2(1.0f, 0.0f);
        });
    }
    SMFFT_DEV void store(float2* stw, int tid) const
    {
        static_for<(K > 0 ? K : 0)>([&](auto KI) {
            constexpr int j = decltype(KI)::value;
            const int i = tid + j * C::THREADS;
            if (i < C::TW_C2C_ENTRIES) {
                float2 x = w[j];
                if (C::DIR) x.y = -x.y;
                stw[i] = x;
            }
        });
    }
};

// ---- tile <-> registers ------------------------------------------------------------------------

// v[m] = tile[fbase + t + m*T]  (natural "column" ownership)
//
// T >= 16: 16 consecutive lanes read one 128-byte row: conflict-free in SW128.
// T < 16 (N <= 128): a half-warp spans 16/T transforms whose rows start at multiples of T, so transforms
// f and f + 8/T see the same swizzle key (row & 7) and collide 2-way.  Those lanes instead fetch
// register i from element i ^ 8 (offset x ^ 8T: a row of the same transform whose key differs in a
// bit ABOVE the column bits the T lanes occupy), and the registers are swapped back with selects:
// 2 SEL per point instead of a second shared-memory wavefront per access.
template <class C, class LY>
struct NaturalAccess {
    static constexpr bool SKEW = C::SKEW_SMALL && (C::T < 16) && std::is_same<LY, LayoutSW128>::value && (C::R == 16);
    static constexpr int D = SKEW ? 8 : 0;
    static SMFFT_DEV int skew(int fbase)
    {
        if constexpr (SKEW)
            return ((fbase >> (C::E + 3 - C::A)) & 1) * (8 * C::T);  // bit (3 - a) of the FFT index -> flip element bit 3
        else
            return 0;
    }
};

template <class C, class LY = typename C::Layout>
SMFFT_DEV void load_natural(float2 (&v)[C::R], const float2* s, int fbase, int t)
{
    using NA = NaturalAccess<C, LY>;
    if constexpr (C::T % 128 == 0 && C::N % 128 == 0 && !std::is_same<LY, LayoutSW256>::value && !std::is_same<LY, LayoutSW128H>::value &&
                  !std::is_same<LY, LayoutSW128R4>::value) {
        const int p0 = LY::phys(fbase + t);  // bits 4..6 do not depend on m
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = plat::lds64(s + p0 + m * C::T);
        });
    } else if constexpr (C::T == 128 && C::N % 256 == 0 && std::is_same<LY, LayoutSW128H>::value) {
        // the key depends on m only through bit 7 of x = t + m*128, i.e. on the parity of m: two bases, constant offsets
        const int pe = LY::phys(fbase + t), po = LY::phys(fbase + t + C::T) - C::T;
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = plat::lds64(s + ((m & 1) ? po : pe) + m * C::T);
        });
    } else if constexpr (NA::SKEW) {
        const int sk = NA::skew(fbase);
        float2 a[C::R];
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            a[m] = plat::lds64(s + LY::phys((fbase + t + m * C::T) ^ sk));
        });
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = sk ? a[m ^ NA::D] : a[m];
        });
    } else {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = plat::lds64(s + LY::phys(fbase + t + m * C::T));
        });
    }
}

template <class C, class LY = typename C::Layout>
SMFFT_DEV void store_natural(const float2 (&v)[C::R], float2* s, int fbase, int t)
{
    using NA = NaturalAccess<C, LY>;
    if constexpr (C::T % 128 == 0 && C::N % 128 == 0 && !std::is_same<LY, LayoutSW256>::value && !std::is_same<LY, LayoutSW128H>::value &&
                  !std::is_same<LY, LayoutSW128R4>::value) {
        const int p0 = LY::phys(fbase + t);
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            plat::sts64(s + p0 + m * C::T, v[m]);
        });
    } else if constexpr (NA::SKEW) {
        const int sk = NA::skew(fbase);
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            const float2 a = sk ? v[m ^ NA::D] : v[m];
            plat::sts64(s + LY::phys((fbase + t + m * C::T) ^ sk), a);
        });
    } else {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            plat::sts64(s + LY::phys(fbase + t + m * C::T), v[m]);
        });
    }
}

// virtual thread id used by the first pass of the no-reorder transform (see header comment)
template <class C>
SMFFT_DEV int noreorder_vid(int t)
{
    if constexpr (C::A >= 3) {
        const int l = t & 7;
        const int rev3 = ((l & 1) << 2) | (l & 2) | ((l >> 2) & 1);
        return t ^ ((rev3 << (C::A - 3)) & ~7 & (C::T - 1));
    } else {
        return t;
    }
}

// v[m] = tile[fbase + brev_e(j + m*T)] = tile[fbase + brev_a(j)*R + brev_b(m)]: one contiguous run
template <class C>
SMFFT_DEV void load_rows_brev(float2 (&v)[C::R], const float2* s, int fbase, int j)
{
    const int row0 = fbase + (int)(plat::brev32((unsigned)j) >> (32 - C::A)) * C::R;
    if constexpr (C::VEC128) {
        static_for<C::R / 2>([&](auto CI) {
            constexpr int c = decltype(CI)::value;
            const float4 q = plat::lds128(s + C::Layout::phys(row0 + 2 * c));
            v[brev_c(2 * c, C::B)] = make_float2(q.x, q.y);
            v[brev_c(2 * c + 1, C::B)] = make_float2(q.z, q.w);
        });
    } else {
        static_for<C::R>([&](auto MI) {
            constexpr int m = decltype(MI)::value;
            v[brev_c(m, C::B)] = plat::lds64(s + C::Layout::phys(row0 + m));
        });
    }
}

// ---- one register pass (+ the autosort exchange that follows it) ---------------------------------

template <class C, int PIDX>
SMFFT_DEV void fft_pass_compute(float2 (&v)[C::R], int vt, const float2* tw)
{
    constexpr int c = C::radix_log2(PIDX), r = 1 << c, U = C::R / r;
    constexpr int NS = 1 << C::ns_log2(PIDX);  // product of the radices already applied
    if constexpr (NS > 1) {
        constexpr int WN = NS * r;
        float2 pw[r];
        make_twiddle_powers<C::DIR, C::TW, WN, r>(pw, vt & (NS - 1), tw + C::tw_offset(PIDX));
        // when NS > T the butterflies of one thread sit in different residue classes mod NS:
        // (vt + u*T) mod NS = vt + (u mod D)*T, which adds the constant factor W_{D r}^{(u mod D) q}
        constexpr int D = NS > C::T ? NS / C::T : 1;
        static_for<U>([&](auto UI) {
            constexpr int u = decltype(UI)::value;
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                if constexpr (q >= 1) {
                    float2 x = cmul(v[u + q * U], pw[q]);
                    if constexpr (D > 1) x = mul_wconst<C::DIR, ((u % D) * q) % (D * r), D * r>(x);
                    v[u + q * U] = x;
                }
            });
        });
    }
    static_for<U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        dft_regs<C::PACK, C::DIR, r, u, U, C::R>(v);
    });
}

template <class C>
struct MirrorC2R;

template <class C, int PIDX, class XL = typename C::XLayout>
SMFFT_DEV void fft_pass_scatter(const float2 (&v)[C::R], float2* s, int fbase, int vt, bool mirror_c2r = false)
{
    constexpr int c = C::radix_log2(PIDX), r = 1 << c, U = C::R / r;
    constexpr int LNS = C::ns_log2(PIDX), NS = 1 << LNS;
    static_for<U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        int j = vt + u * C::T;
        if constexpr (PIDX == 0 && U >= 2) {
            if (mirror_c2r) j = MirrorC2R<C>::vthread(vt, u);  // mirrored ownership of the C2R first pass
        }
        const int xb = fbase + ((j >> LNS) << (LNS + c)) + (j & (NS - 1));
        if constexpr (NS == 1 && C::VEC128) {
            // r contiguous outputs per butterfly: 128-bit stores
            static_for<r / 2>([&](auto QI) {
                constexpr int q = 2 * decltype(QI)::value;
                const float2 lo = v[u + q * U], hi = v[u + (q + 1) * U];
                plat::sts128(s + XL::phys(xb + q), make_float4(lo.x, lo.y, hi.x, hi.y));
            });
        } else if constexpr (NS == 8 && r == 16 && std::is_same<XL, LayoutSW128H>::value) {
            // x = xb + 8q without carries (xb holds bits 0..2 and 7.., 8q bits 3..6): the SW128H key of x is
            // ((q >> 1) & 7) ^ (bit 7 of xb << 2), so phys(x) = (xb ^ (bit7 << 3)) ^ (8q ^ (((q >> 1) & 7) << 1)): one XOR per store
            const int bq = xb ^ (((xb >> 7) & 1) << 3);
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                plat::sts64(s + (bq ^ ((8 * q) ^ (((q >> 1) & 7) << 1))), v[u + q * U]);
            });
        } else if constexpr (NS % 256 == 0) {
            const int p0 = XL::phys(xb);  // q*NS leaves bits 0..7 alone (SW128 keys on bits 4..6, SW256 on 5..7)
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                plat::sts64(s + p0 + q * NS, v[u + q * U]);
            });
        } else {
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                plat::sts64(s + XL::phys(xb + q * NS), v[u + q * U]);
            });
        }
    });
}

struct NoHook {
    static constexpr int PASS = 0;
    static constexpr bool TAIL = false;
    SMFFT_DEV void operator()() const {}
    SMFFT_DEV void tail() const {}
};
// a callable to run once per tile, right after the first barrier that follows pass PASS (clamped to
// the last exchange of the plan): the kernels use it to time the refill of the previous tile buffer
template <int PASS_, class F>
struct HookAt {
    static constexpr int PASS = PASS_;
    static constexpr bool TAIL = false;
    F f;
    SMFFT_DEV void operator()() const { f(); }
    SMFFT_DEV void tail() const {}
};
template <int PASS_, class F>
SMFFT_DEV HookAt<PASS_, F> hook_at(F f)
{
    return HookAt<PASS_, F>{f};
}
// a TAIL callable: every thread runs it once per tile right after ITS last read of the tile buffer (the operands of the
// last pass, i.e. the final exchange); the results leave from registers.  Behind a block barrier inside the callable the
// buffer is dead, so a TMA load into the very buffer the transform ran in may start there (single-buffer kernels: 16384
// points, where two 128 KB tiles do not fit one SM) -- the callable also owns the generic -> async proxy fence that needs.
template <class G>
struct HookTail {
    static constexpr int PASS = -2;  // no regular hook
    static constexpr bool TAIL = true;
    G g;
    SMFFT_DEV void operator()() const {}
    SMFFT_DEV void tail() const { g(); }
};
template <class G>
SMFFT_DEV HookTail<G> hook_tail(G g)
{
    return HookTail<G>{g};
}

// layout of the exchange that follows pass PIDX
template <class C, int PIDX>
struct ExchangeLayout {
    static constexpr bool SW = std::is_same<typename C::Layout, LayoutSW128>::value;
    static constexpr bool H = C::ns_log2(PIDX) == 3 && std::is_same<typename C::XLayout, LayoutSW128>::value;
    static constexpr bool Q = C::REV && PIDX == 0 && C::radix_log2(0) == 3 && std::is_same<typename C::XLayout, LayoutSW128>::value;
    static constexpr bool P2 = C::REV && SW && PIDX == 0 && C::radix_log2(0) == 2;   // radix-4 first pass
    static constexpr bool R4 = C::REV && SW && C::ns_log2(PIDX) == 2 && PIDX >= 1;   // the Ns = 4 exchange after it
    using type = typename std::conditional<P2, LayoutSW128P, typename std::conditional<R4, LayoutSW128R4,
                 typename std::conditional<H, LayoutSW128H, typename std::conditional<Q, LayoutSW128Q, typename C::XLayout>::type>::type>::type>::type;
};
// the layout the LAST pass reads from = where the in-place result must not be written without a barrier
template <class C>
struct LastExchangeSameAsTile {
    static constexpr bool value = C::P < 2 || std::is_same<typename ExchangeLayout<C, (C::P >= 2 ? C::P - 2 : 0)>::type, typename C::Layout>::value;
};

// ---- C2R with mirrored ownership in the FIRST pass (reversed plan) ------------------------------------------------
// The mirror image of MirrorR2C: with the small radix first (r0 = R/2, U = 2 butterflies per thread) the first pass reads
// x = j + q 2T for two virtual threads j; choosing j = t and j' = 2T - t (thread 0: 0 and T) puts Y[k] and Y[N-k] in one
// thread, so the inverse real pass runs in pair form on registers: 16 instead of 32 LDS per thread and 12 instead of 20
// floating-point instructions per pair.  Pass 0 has no twiddles (Ns = 1); only its scatter uses j' for the second butterfly.
template <class C>
struct MirrorC2R {
    static constexpr int r = 1 << C::radix_log2(0);
    static constexpr int U = C::R / r;    // butterflies per thread in the first pass = U/2 mirror pairs
    static constexpr int NS2 = C::N / r;  // = U T virtual threads
    static constexpr bool OK = !C::DUAL && C::REV && C::P >= 2 && (C::R == 16 || C::R == 32) && U >= 2 && r >= 4 && C::T >= 16 && C::VEC128 &&
                               C::REORDER == 1 && std::is_same<typename C::Layout, LayoutSW128>::value;
    static SMFFT_DEV int vthread(int t, int u)
    {
        const int a = t + (u >> 1) * C::T;
        return (u & 1) ? (a == 0 ? NS2 / 2 : NS2 - a) : a;
    }
};

// one pair of the inverse real pass: A = Y[k], Bv = Y[N-k], Wh = exp(+2 pi i k / 2N) / 2  ->  zk = Z[k], zn = Z[N-k]
// (real_combine<1> evaluated for k and for N-k, sharing the sums: W^{N-k} = -conj W^k)
template <int PACK>
SMFFT_DEV void c2r_pair(float2 A, float2 Bv, float2 Wh, float2& zk, float2& zn)
{
    const float2 sm = cadd_p<PACK>(A, Bv), df = csub_p<PACK>(A, Bv);
    const float sx = sm.x, sy = sm.y, dx = df.x, dy = df.y;
    const float p = Wh.x * sy + Wh.y * dx, q = Wh.y * sy - Wh.x * dx;
    zk = make_float2(0.5f * sx - p, 0.5f * dy - q);
    zn = make_float2(0.5f * sx + p, -0.5f * dy - q);
}

// v[u + qU] = Y[vthread(t, u) + q Ns] from the (read-only, SW128) tile, then the inverse real pass on the pairs
// (pair i: butterfly 2i output q  <->  butterfly 2i+1 output r-1-q; the slot a = 0 pairs within itself, see r2c_tail_mirror)
template <class C>
SMFFT_DEV void c2r_head_mirror(float2 (&v)[C::R], const float2* s, int fbase, int t, const float2* tw)
{
    using M = MirrorC2R<C>;
    constexpr int r = M::r, U = M::U;
    static_assert(4 * r <= 64 && 2 * C::R <= 64, "constant twiddles of the mirrored real pass come from the W_64 table");
    static_for<U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        const int j = M::vthread(t, u);
        if constexpr (M::NS2 % 128 == 0) {
            const int p = C::Layout::phys(fbase + j);  // q*Ns moves whole groups of eight rows: the swizzle key stays
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                v[u + q * U] = plat::lds64(s + p + q * M::NS2);
            });
        } else {
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                v[u + q * U] = plat::lds64(s + C::Layout::phys(fbase + j + q * M::NS2));
            });
        }
    });
    float2 wt;  // exp(+2 pi i t / 2N) / 2
    if constexpr (C::TW == TW_LUT) {
        wt = plat::lds64(tw + C::TW_C2C_ENTRIES + t);
    } else {
        wt = tw_mufu<1, 2 * C::N>(t);
        wt.x *= 0.5f;
        wt.y *= 0.5f;
    }
    auto general = [&](auto II) {
        constexpr int i = decltype(II)::value;
        static_for<r>([&](auto QI) {
            constexpr int q = decltype(QI)::value;
            constexpr int ma = 2 * i + q * U, mb = 2 * i + 1 + (r - 1 - q) * U;
            c2r_pair<C::PACK>(v[ma], v[mb], mul_wconst<1, i + q * U, 2 * C::R>(wt), v[ma], v[mb]);
        });
    };
    static_for<U / 2>([&](auto II) {
        if constexpr (decltype(II)::value >= 1) general(II);
    });
    if (t != 0) {
        general(std::integral_constant<int, 0>{});
    } else {
        const float2 y0 = v[0], ym = v[(r / 2) * U];
        v[0] = make_float2(0.5f * (y0.x + y0.y), 0.5f * (y0.x - y0.y));  // bin 0 un-packed (RC:280-286)
        v[(r / 2) * U] = make_float2(ym.x, -ym.y);                       // k = N/2 is its own partner
        static_for<r / 2>([&](auto QI) {
            constexpr int q = decltype(QI)::value;
            if constexpr (q >= 1)
                c2r_pair<C::PACK>(v[q * U], v[(r - q) * U], mul_wconst<1, q, 2 * r>(make_float2(0.5f, 0.0f)), v[q * U], v[(r - q) * U]);
        });
        static_for<r / 2>([&](auto QI) {
            constexpr int q = decltype(QI)::value;  // k = Ns/2 + q Ns
            c2r_pair<C::PACK>(v[1 + q * U], v[1 + (r - 1 - q) * U], mul_wconst<1, 1 + 2 * q, 4 * r>(make_float2(0.5f, 0.0f)), v[1 + q * U],
                              v[1 + (r - 1 - q) * U]);
        });
    }
}

// ---- R2C with mirrored ownership in the last pass (RC/FFT-GPU-32bit-Stockham.cu:269-344 without its exchange) ----
// The real pass needs Z[k] and Z[N-k] together.  When the last pass runs U = 2 butterflies per thread (radix r = R/2,
// Ns = N/r = 2T virtual threads), the thread may pick WHICH two virtual threads it computes: j = t and the mirror
// j' = Ns - t (thread 0: j = 0 and j' = T, both their own mirrors).  Its outputs are then k = j + q Ns and
// N - k = j' + (r-1-q) Ns: every pair (k, N-k) sits in ONE thread and the real pass needs no exchange and no barrier
// -- 8 STS + 8 LDS + 1 barrier less per 16 points than r2c_tail_regs (the kernels are bound by shared-memory
// wavefronts, profiles/r01_ncu_summary_z.md).  The last exchange uses a LINEAR layout: its scatter writes and the
// ascending reads are whole 128-byte rows, and the descending mirror reads j' = Ns - t are conflict-free only there
// (lane 0 wraps into column 0 of the next row, the one bank pair the other 15 lanes leave free).
// Twiddles of the mirrored butterfly: W_N^{(Ns - t) q} = W_r^q conj(W_N^{t q});  thread 0: W_N^{T q} = W_{2r}^q.
// With U > 2 butterflies per thread the same holds for U/2 pairs (a_i = t + iT, Ns - a_i).
template <class C>
struct MirrorR2C {
    static constexpr int PL = C::P - 1;
    static constexpr int r = 1 << C::radix_log2(PL);
    static constexpr int U = C::R / r;   // butterflies per thread in the last pass = U/2 mirror pairs
    static constexpr int NS = C::N / r;  // = U T virtual threads
    // general U (even): pair i takes a_i = t + i T and its mirror Ns - a_i (the one self-mirrored slot a = 0 takes Ns/2)
    static constexpr bool OK = !C::DUAL && C::P >= 3 && (C::R == 16 || C::R == 32) && U >= 2 && r >= 2 && C::T >= 16 && C::VEC128 &&
                               C::REORDER == 1 && (1 << C::ns_log2(C::P - 2)) >= 16 && std::is_same<typename C::Layout, LayoutSW128>::value;
    static SMFFT_DEV int vthread(int t, int u)
    {
        const int a = t + (u >> 1) * C::T;
        return (u & 1) ? (a == 0 ? NS / 2 : NS - a) : a;
    }
    // index (within the transform) of register m = u + q U
    static SMFFT_DEV int index(int t, int m) { return vthread(t, m % U) + (m / U) * NS; }
};

template <class C>
SMFFT_DEV void load_mirror(float2 (&v)[C::R], const float2* s, int fbase, int t)
{
    using M = MirrorR2C<C>;
    static_for<M::U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        const float2* a = s + fbase + M::vthread(t, u);
        static_for<M::r>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            v[u + q * M::U] = plat::lds64(a + q * M::NS);
        });
    });
}

template <class C>
SMFFT_DEV void fft_pass_compute_mirror(float2 (&v)[C::R], int t, const float2* tw)
{
    using M = MirrorR2C<C>;
    constexpr int r = M::r, U = M::U, R = C::R;
    float2 pw[r];
    make_twiddle_powers<C::DIR, C::TW, C::N, r>(pw, t, tw + C::tw_offset(M::PL));
    static_for<U / 2>([&](auto II) {
        constexpr int i = decltype(II)::value;
        static_for<r>([&](auto QI) {
            constexpr int q = decltype(QI)::value;
            if constexpr (q >= 1) {
                // a_i = t + i T:  W_N^{a q} = W_N^{t q} W_R^{i q}
                v[2 * i + q * U] = mul_wconst<C::DIR, (i * q) % R, R>(cmul(v[2 * i + q * U], pw[q]));
                // mirror Ns - a:  W_r^q conj(W_N^{a q}) = conj(W_N^{t q}) W_R^{qU - iq};  a = 0 (thread 0, pair 0) owns Ns/2
                // instead: W_N^{(Ns/2) q} = W_{2r}^q = W_r^q conj(W_{2r}^q)
                float2 wc = make_float2(pw[q].x, -pw[q].y);
                if constexpr (i == 0) {
                    constexpr float cx = cos64(q * (64 / (2 * r))), sy = C::DIR ? sin64(q * (64 / (2 * r))) : -sin64(q * (64 / (2 * r)));
                    if (t == 0) wc = make_float2(cx, -sy);
                }
                v[2 * i + 1 + q * U] = mul_wconst<C::DIR, (((q * U - i * q) % R) + R) % R, R>(cmul(v[2 * i + 1 + q * U], wc));
            }
        });
    });
    static_for<U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        dft_regs<C::PACK, C::DIR, r, u, U, C::R>(v);
    });
}

// one pair of the real pass: A = Z[k], Bv = Z[N-k], Wh = W_{2N}^k / 2  ->  lo = X[k], hi = X[N-k] (see r2c_tail_regs)
template <int PACK>
SMFFT_DEV void r2c_pair(float2 A, float2 Bv, float2 Wh, float2& lo, float2& hi)
{
    const float2 sm = cadd_p<PACK>(A, Bv), df = csub_p<PACK>(A, Bv);  // (sx, dy), (dx, sy)
    const float sx = sm.x, sy = df.y, dx = df.x, dy = sm.y;
    const float px = Wh.x * dy + Wh.y * dx, py = Wh.y * dy - Wh.x * dx;
    lo = make_float2(0.5f * sx + px, 0.5f * sy + py);
    hi = make_float2(0.5f * sx - px, py - 0.5f * sy);
}

// real pass on the mirrored ownership, registers only.  On return register m holds X[MirrorR2C::index(t, m)].
// Pair i, butterfly 2i (a = t + iT), output q:  k = a + q Ns  <->  N - k = (Ns - a) + (r-1-q) Ns = butterfly 2i+1, output r-1-q;
// W_{2N}^k = W_{2N}^t W_{2R}^{i + qU}.  The slot a = 0 (thread 0, pair 0) pairs within itself: k = q Ns <-> (r-q) Ns, bin 0
// packs (X[0], X[N]), k = N/2 is its own partner; its sibling Ns/2 pairs k = Ns/2 + q Ns <-> Ns/2 + (r-1-q) Ns.
template <class C>
SMFFT_DEV void r2c_tail_mirror(float2 (&v)[C::R], const float2* tw)
{
    using M = MirrorR2C<C>;
    constexpr int r = M::r, U = M::U;
    static_assert(4 * r <= 64 && 2 * C::R <= 64, "constant twiddles of the mirrored real pass come from the W_64 table");
    const int t = plat::tid() & (C::T - 1);
    float2 wt;  // W_{2N}^t / 2
    if constexpr (C::TW == TW_LUT) {
        wt = plat::lds64(tw + C::TW_C2C_ENTRIES + t);
    } else {
        wt = tw_mufu<0, 2 * C::N>(t);
        wt.x *= 0.5f;
        wt.y *= 0.5f;
    }
    auto general = [&](auto II) {
        constexpr int i = decltype(II)::value;
        static_for<r>([&](auto QI) {
            constexpr int q = decltype(QI)::value;
            constexpr int ma = 2 * i + q * U, mb = 2 * i + 1 + (r - 1 - q) * U;
            r2c_pair<C::PACK>(v[ma], v[mb], mul_wconst<0, i + q * U, 2 * C::R>(wt), v[ma], v[mb]);
        });
    };
    static_for<U / 2>([&](auto II) {
        if constexpr (decltype(II)::value >= 1) general(II);
    });
    if (t != 0) {
        general(std::integral_constant<int, 0>{});
    } else {
        const float2 z0 = v[0], zm = v[(r / 2) * U];
        v[0] = make_float2(z0.x + z0.y, z0.x - z0.y);
        v[(r / 2) * U] = make_float2(zm.x, -zm.y);
        static_for<r / 2>([&](auto QI) {
            constexpr int q = decltype(QI)::value;
            if constexpr (q >= 1)
                r2c_pair<C::PACK>(v[q * U], v[(r - q) * U], mul_wconst<0, q, 2 * r>(make_float2(0.5f, 0.0f)), v[q * U], v[(r - q) * U]);
        });
        static_for<r / 2>([&](auto QI) {
            constexpr int q = decltype(QI)::value;  // k = Ns/2 + q Ns:  W_{2N}^k = W_{4r}^{1 + 2q}
            r2c_pair<C::PACK>(v[1 + q * U], v[1 + (r - 1 - q) * U], mul_wconst<0, 1 + 2 * q, 4 * r>(make_float2(0.5f, 0.0f)), v[1 + q * U],
                              v[1 + (r - 1 - q) * U]);
        });
    }
}

// hook() runs once per tile, right after the first barrier following pass Hook::PASS: from the first
// barrier of a tile onwards every thread of the CTA has finished ALL work of the previous tile, so the
// previous tile's buffer may be refilled; a later pass shortens the time the refill is in flight.
template <class C, int PIDX, int XF = 0, class Hook = NoHook>
SMFFT_DEV void run_passes(float2 (&v)[C::R], float2* s, int fbase, int vt, int t, const float2* tw, Hook&& hook)
{
    constexpr bool MIRROR = XF == 1 /* XF_R2C */ && MirrorR2C<C>::OK;
    if constexpr (MIRROR && PIDX == C::P - 1)
        fft_pass_compute_mirror<C>(v, t, tw);
    else
        fft_pass_compute<C, PIDX>(v, vt, tw);
    if constexpr (C::XSHFL && PIDX == 0) {
        // Plan [16, T], T = 2 or 4 lanes per transform: output q of lane j belongs to lane q mod T, register (16/T) j + q div T.
        // Per group i of T registers {v[iT + c]} that is a T x T transposition with the lanes, then a renaming.
        static_assert(C::P == 2 && (C::T == 2 || C::T == 4) && C::R == 16, "shuffle exchange: plan [16, T]");
        const int lane = plat::tid() & 31;
        float2 w[C::R];
        static_for<C::R / C::T>([&](auto II) {
            constexpr int i = decltype(II)::value;
            if constexpr (C::T == 4) {
                float2 g[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
                wf::xchg4<0, 1>(g, lane);
                static_for<4>([&](auto JI) { w[(C::R / C::T) * decltype(JI)::value + i] = g[decltype(JI)::value]; });
            } else {
                const bool b = lane & 1;
                const float2 send = b ? v[2 * i] : v[2 * i + 1];
                const float2 got = wf::shfl_xor2(send, 1);
                w[i] = b ? got : v[2 * i];                      // from lane 0: its v[2i + t]
                w[(C::R / C::T) + i] = b ? v[2 * i + 1] : got;  // from lane 1
            }
        });
        static_for<C::R>([&](auto MI) { v[decltype(MI)::value] = w[decltype(MI)::value]; });
        // no barrier separates this tile's first read from its final write any more, and with fft_reorder = 0 a lane's
        // result columns overlap the ROW its neighbour read: order them (all lanes of a transform share a warp)
        plat::sync_warp();
        run_passes<C, PIDX + 1, XF>(v, s, fbase, t, t, tw, hook);
    } else if constexpr (PIDX + 1 < C::P) {
        plat::sync_block();  // every thread has finished reading the previous state of the tile
        if constexpr (PIDX == (std::remove_reference<Hook>::type::PASS < C::P - 2 ? std::remove_reference<Hook>::type::PASS : C::P - 2)) hook();
        using XL = typename ExchangeLayout<C, PIDX>::type;
        if constexpr (MIRROR && PIDX + 1 == C::P - 1) {
            fft_pass_scatter<C, PIDX, LayoutLinear>(v, s, fbase, vt);
            plat::sync_block();
            load_mirror<C>(v, s, fbase, t);
        } else {
            if constexpr (XF == 2 /* XF_C2R */ && MirrorC2R<C>::OK && PIDX == 0)
                fft_pass_scatter<C, PIDX, XL>(v, s, fbase, vt, true);
            else
                fft_pass_scatter<C, PIDX, XL>(v, s, fbase, vt);
            plat::sync_block();
            load_natural<C, XL>(v, s, fbase, t);
            if constexpr (std::remove_reference<Hook>::type::TAIL && PIDX + 1 == C::P - 1) hook.tail();  // this thread's last read of the tile buffer is done
        }
        run_passes<C, PIDX + 1, XF>(v, s, fbase, t, t, tw, hook);
    }
}

enum { XF_C2C = 0, XF_R2C = 1, XF_C2R = 2 };  // what one tile computes (kernels::MODE_* use the same values)

// ---- R2C / C2R in registers (RC/FFT-GPU-32bit-Stockham.cu:269-344) ---------
// N = C::N complex points hold one real transform of length 2N, z[n] = x[2n] + i x[2n+1].
//   forward : X[k] = H1 + W H2,  H1 = (A + conj B)/2, H2 = -i (A - conj B)/2,  A = Z[k], B = Z[N-k],
//             W = exp(-2 pi i k / 2N), k = 1..N-1;  bin 0 packs (X[0], X[N]) = (Z0.re + Z0.im, Z0.re - Z0.im)
//   inverse : the same with conjugated constants, bin 0 un-packed first
// The reference does this as a separate shared-memory pass over pairs (k, N-k).  Here a thread keeps
// its own 16 values in registers, reads only the 16 partners B from shared memory, and evaluates the
// formula for its own k (every pair is evaluated from both ends: more FMAs, half the shared traffic,
// no extra write-back pass).  The twiddle is one table/MUFU value W^t times the constant W_{2R}^m.
// Wh = W/2 (the table stores it pre-scaled).  With s = A + B, d = A - B (componentwise):
//   forward: X = ( s.x/2 + Wh.x s.y + Wh.y d.x ,  d.y/2 - Wh.x d.x + Wh.y s.y )
//   inverse: Z = ( s.x/2 - Wh.x s.y - Wh.y d.x ,  d.y/2 + Wh.x d.x - Wh.y s.y )   (Wh already conjugated)
// -- the reference's H1 + W H2 (RC:292-307) with the constants folded in: 4 adds + 6 FMAs.
template <int INVERSE, int PACK = 0>
SMFFT_DEV float2 real_combine(float2 A, float2 B, float2 Wh)
{
    const float2 sm = cadd_p<PACK>(A, B), df = csub_p<PACK>(A, B);
    const float sx = sm.x, sy = sm.y, dx = df.x, dy = df.y;
    float2 o;
    if constexpr (!INVERSE) {
        o.x = 0.5f * sx + (Wh.x * sy + Wh.y * dx);
        o.y = 0.5f * dy + (Wh.y * sy - Wh.x * dx);
    } else {
        o.x = 0.5f * sx - (Wh.x * sy + Wh.y * dx);
        o.y = 0.5f * dy - (Wh.y * sy - Wh.x * dx);
    }
    return o;
}

// v[m] holds element k = t + m*T of the tile state `s` (natural order, entry layout); on return
// v[m] = combined value for k.  `s` is only read.
template <class C, int INVERSE>
SMFFT_DEV void real_pass_regs(float2 (&v)[C::R], const float2* s, int fbase, int t, const float2* tw)
{
    static_assert(2 * C::R <= 64, "constant twiddles W_{2R}^m come from the W_64 table");
    float2 wt;  // W_{2N}^t / 2
    if constexpr (C::TW == TW_LUT) {
        wt = plat::lds64(tw + C::TW_C2C_ENTRIES + t);
    } else {
        wt = tw_mufu<INVERSE, 2 * C::N>(t);
        wt.x *= 0.5f;
        wt.y *= 0.5f;
    }
    const int xp = fbase + C::N - t;  // partner of k = t + m*T is N - k = (N - t) - m*T
    // W^{t+mT}/2 = wt * W_{2R}^m; the upper half of the m range is the lower half times a quarter turn (free)
    float2 wm[C::R];
    static_for<C::R / 2>([&](auto MI) {
        constexpr int m = decltype(MI)::value;
        wm[m] = mul_wconst<INVERSE, m, 2 * C::R>(wt);
        wm[m + C::R / 2] = mul_wconst<INVERSE, 1, 4>(wm[m]);
    });
    static_for<C::R>([&](auto MI) {
        constexpr int m = decltype(MI)::value;
        if constexpr (m == 0) {
            // t == 0 owns bin 0, which has no partner: (X[0], X[N]) packed / un-packed (RC:280-286, 332-340)
            const float2 Bv = plat::lds64(s + C::Layout::phys(t == 0 ? fbase : xp));
            float2 out = real_combine<INVERSE, C::PACK>(v[0], Bv, wm[0]);
            if (t == 0) {
                const float sc = INVERSE ? 0.5f : 1.0f;
                out = make_float2(sc * (v[0].x + v[0].y), sc * (v[0].x - v[0].y));
            }
            v[0] = out;
        } else {
            const float2 Bv = plat::lds64(s + C::Layout::phys(xp - m * C::T));
            v[m] = real_combine<INVERSE, C::PACK>(v[m], Bv, wm[m]);
        }
    });
}

// load (+ C2R pre-pass) + all FFT passes; the result is left in registers: v[m] = X[t + m*T] of FFT (tid >> A)
template <class C, int XF, class Hook>
SMFFT_DEV void block_fft_regs(float2 (&v)[C::R], float2* s, const float2* tw, Hook&& hook)
{
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    const int fbase = (tid >> C::A) << C::E;
    int vt = t;
    if constexpr (XF == XF_C2R && MirrorC2R<C>::OK) {
        c2r_head_mirror<C>(v, s, fbase, t, tw);  // tile is read-only here: no barrier
    } else if constexpr (C::REORDER) {
        load_natural<C>(v, s, fbase, t);
        if constexpr (XF == XF_C2R) real_pass_regs<C, 1>(v, s, fbase, t, tw);  // tile is read-only here: no barrier
    } else {
        vt = noreorder_vid<C>(t);
        load_rows_brev<C>(v, s, fbase, vt);
    }
    run_passes<C, 0, XF>(v, s, fbase, vt, t, tw, hook);
}

// R2C tail, pair form.  On entry v[m] = Z[t + m*T].  Thread t evaluates the R/2 pairs (k, N-k), k = t + i*T < N/2:
// A = Z[k] is its own register i, B = Z[N-k] belongs to thread T-t, so only the UPPER half of every thread's
// values goes through the tile (R/2 STS + R/2 LDS per thread instead of a full write and read), and the two
// results share H1 and W*H2:  X[k] = H1 + W H2,  X[N-k] = conj(H1 - W H2)  (RC:292-307 evaluates the same pair).
// With s = A + conj B, d = A - conj B, Wh = W/2:  P = Wh * (d.y, -d.x);  X[k] = s/2 + P;  X[N-k] = conj(s/2 - P)
// -- 4 adds + 8 FMA-class ops per pair.  Thread 0 has no partner for k = 0: it packs bin 0 = (X[0], X[N]) and
// emits the self-paired bin N/2 = conj(Z[N/2]) (its own register R/2) instead.
// On return v[i] = X[t + i*T] and v[R/2 + i] = X[r2c_hi_index(t, i)], i < R/2.
template <class C>
SMFFT_DEV int r2c_hi_index(int t, int i)
{
    return (i == 0 && t == 0) ? C::N / 2 : C::N - t - i * C::T;
}
// tile index (within the FFT) of register m after r2c_tail_regs
template <class C>
SMFFT_DEV int r2c_out_index(int t, int m)
{
    if constexpr (MirrorR2C<C>::OK)
        return MirrorR2C<C>::index(t, m);
    else
        return m < C::R / 2 ? t + m * C::T : r2c_hi_index<C>(t, m - C::R / 2);
}

template <class C>
SMFFT_DEV void r2c_tail_regs(float2 (&v)[C::R], float2* s, const float2* tw)
{
    static_assert(2 * C::R <= 64, "constant twiddles W_{2R}^i come from the W_64 table");
    if constexpr (MirrorR2C<C>::OK) {  // every pair already sits in one thread (run_passes<.., XF_R2C>): no exchange
        r2c_tail_mirror<C>(v, tw);
        return;
    }
    constexpr int H = C::R / 2;
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    const int fbase = (tid >> C::A) << C::E;
    // same layout: these are slots this thread read in the last pass, no barrier needed before.  Otherwise a barrier is
    // due anyway, and behind it the scratch may use ANY layout: linear, where the ascending writes and the descending
    // partner reads are both conflict-free (a descending run collides 2-way with its wrapped lane under SW128)
    // (paying an extra barrier for the linear scratch in the same-layout kernels loses 0.6-0.8 %, measured)
    constexpr bool DUE = (!C::SAME_LAYOUT || !LastExchangeSameAsTile<C>::value) && C::P > 1;
    constexpr bool FREE = DUE && C::T >= 16;
    using TL = typename std::conditional<FREE, LayoutLinear, typename C::Layout>::type;
    if constexpr (DUE) plat::sync_block();
    static_for<H>([&](auto II) {
        constexpr int m = H + decltype(II)::value;
        plat::sts64(s + TL::phys(fbase + t + m * C::T), v[m]);
    });
    float2 wt;  // W_{2N}^t / 2
    if constexpr (C::TW == TW_LUT) {
        wt = plat::lds64(tw + C::TW_C2C_ENTRIES + t);
    } else {
        wt = tw_mufu<0, 2 * C::N>(t);
        wt.x *= 0.5f;
        wt.y *= 0.5f;
    }
    const float2 zmid = v[H];
    plat::sync_block();
    static_for<H>([&](auto II) {
        constexpr int i = decltype(II)::value;
        const float2 Wh = mul_wconst<0, i, 2 * C::R>(wt);  // W_{2N}^{t + i T} / 2
        const float2 A = v[i];
        const int xb = (i == 0 && t == 0) ? fbase + C::N / 2 : fbase + C::N - t - i * C::T;
        const float2 Bv = plat::lds64(s + TL::phys(xb));
        float2 lo, hi;
        r2c_pair<C::PACK>(A, Bv, Wh, lo, hi);
        if constexpr (i == 0) {
            if (t == 0) {
                lo = make_float2(A.x + A.y, A.x - A.y);
                hi = make_float2(zmid.x, -zmid.y);
            }
        }
        v[i] = lo;
        v[H + i] = hi;
    });
}

// registers -> tile after the transform: natural columns, or the pair ownership left by r2c_tail_regs
template <class C, int XF>
SMFFT_DEV void store_result(const float2 (&v)[C::R], float2* s, int fbase, int t)
{
    if constexpr (XF == XF_R2C && MirrorR2C<C>::OK && MirrorR2C<C>::NS % 128 == 0) {
        using M = MirrorR2C<C>;  // q*Ns moves whole groups of eight rows: one swizzled base per butterfly, constant offsets
        static_for<M::U>([&](auto UI) {
            constexpr int u = decltype(UI)::value;
            const int p = C::Layout::phys(fbase + M::vthread(t, u));
            static_for<M::r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                plat::sts64(s + p + q * M::NS, v[u + q * M::U]);
            });
        });
    } else if constexpr (XF == XF_R2C) {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            plat::sts64(s + C::Layout::phys(fbase + r2c_out_index<C>(t, m)), v[m]);
        });
    } else {
        store_natural<C>(v, s, fbase, t);
    }
}

// In-place transform of all F transforms of the tile (XF_C2C / XF_R2C / XF_C2R).  Contract: the tile
// is visible to the whole CTA on entry; on return the caller must synchronise before other threads
// (or the async proxy) read the tile.
template <class C, int XF, class Hook>
SMFFT_DEV void dual_fft_tile(float2* s, const float2* tw, Hook&& hook);
template <class C, int XF, class Hook>
SMFFT_DEV void dual_fft_tile_to_global(float2* s, const float2* tw, float2* __restrict__ g, long long valid, Hook&& hook);

template <class C, int XF = XF_C2C, class Hook = NoHook>
SMFFT_DEV void block_fft_tile(float2* s, const float2* tw, Hook&& hook = Hook{})
{
    if constexpr (C::DUAL) {
        dual_fft_tile<C, XF>(s, tw, hook);
    } else {
        float2 v[C::R];
        block_fft_regs<C, XF>(v, s, tw, hook);
        const int tid = plat::tid();
        if constexpr (XF == XF_R2C) {
            r2c_tail_regs<C>(v, s, tw);
            plat::sync_block();  // every partner has been read before the packed spectrum overwrites Z
        } else {
            // with distinct entry/exchange layouts the final slots are not the ones this thread just read
            if constexpr ((!C::SAME_LAYOUT || !LastExchangeSameAsTile<C>::value) && C::P > 1) plat::sync_block();
        }
        store_result<C, XF>(v, s, (tid >> C::A) << C::E, tid & (C::T - 1));
    }
}

// registers -> global memory (coalesced 8-byte stores: consecutive threads own consecutive points; after the
// R2C tail the upper half of the registers holds the descending run of the pair partners).
// g = start of this tile in the output, valid = points of the batch left from there.
template <class C, int XF>
SMFFT_DEV void store_global_result(const float2 (&v)[C::R], float2* __restrict__ g, long long valid)
{
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    const int fbase = (tid >> C::A) << C::E;
    auto index = [&](auto M) {
        constexpr int m = decltype(M)::value;
        if constexpr (XF == XF_R2C)
            return fbase + r2c_out_index<C>(t, m);
        else
            return fbase + t + m * C::T;
    };
    if (valid >= C::L) {
        static_for<C::R>([&](auto M) { plat::stg64_stream(g + index(M), v[decltype(M)::value]); });
    } else {
        static_for<C::R>([&](auto M) {
            const int x = index(M);
            if (x < valid) plat::stg64_stream(g + x, v[decltype(M)::value]);
        });
    }
}

// Same transform, result written straight from registers to global memory.
template <class C, int XF, class Hook>
SMFFT_DEV void block_fft_tile_to_global(float2* s, const float2* tw, float2* __restrict__ g, long long valid, Hook&& hook)
{
    if constexpr (C::DUAL) {
        dual_fft_tile_to_global<C, XF>(s, tw, g, valid, hook);
    } else {
        float2 v[C::R];
        block_fft_regs<C, XF>(v, s, tw, hook);
        if constexpr (XF == XF_R2C) r2c_tail_regs<C>(v, s, tw);
        store_global_result<C, XF>(v, g, valid);
    }
}

// ---- register-direct input (kernels IO_REG): global -> registers -> passes -> global -----------------
// Natural-order transforms only: thread t of an FFT owns x = t + m*T, so every warp-level load is one
// contiguous run (>= 32 bytes per FFT).  The tile buffer `s` then carries only the exchanges between passes:
// 32 instead of 64 bytes of shared-memory traffic per point for a two-pass plan.
template <class C>
SMFFT_DEV void load_global_natural(float2 (&v)[C::R], const float2* __restrict__ g, long long valid)
{
    const int tid = plat::tid();
    const int x0 = ((tid >> C::A) << C::E) + (tid & (C::T - 1));
    if (valid >= C::L) {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = plat::ldg64_stream(g + x0 + m * C::T);
        });
    } else {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = (x0 + m * C::T < valid) ? plat::ldg64_stream(g + x0 + m * C::T) : make_float2(0.0f, 0.0f);
        });
    }
}

// v = the thread's points of this tile (load_global_natural).  `s` may still be read by slower threads working
// on the previous tile: the first barrier of run_passes orders that.
template <class C, int XF, class Hook = NoHook>
SMFFT_DEV void block_fft_preloaded_to_global(float2 (&v)[C::R], float2* s, const float2* tw, float2* __restrict__ g,
                                             long long valid, Hook&& hook = Hook{})
{
    static_assert(C::REORDER == 1 && XF != XF_C2R, "register-direct input: natural-order C2C and R2C");
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    const int fbase = (tid >> C::A) << C::E;
    run_passes<C, 0, XF>(v, s, fbase, t, t, tw, hook);
    if constexpr (XF == XF_R2C) r2c_tail_regs<C>(v, s, tw);
    store_global_result<C, XF>(v, g, valid);
}

// ---- R2C / C2R pair pass, shared-memory form (used by include/smfft/compat.cuh, 4 points/thread) ----
// M = C::N complex points hold one real transform of length 2M.  Each FFT has M/2 pairs (k, M-k),
// k = 1..M/2, spread over its T threads (R/2 pairs per thread), plus bin 0 on thread 0.
// INVERSE = 0: call after the forward C2C;  INVERSE = 1: call before the inverse C2C.
template <class C, int INVERSE>
SMFFT_DEV void r2c_pair_pass_tile(float2* s, const float2* tw)
{
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    const int fbase = (tid >> C::A) << C::E;
    constexpr int M = C::N;
    constexpr float hx = INVERSE ? -0.5f : 0.5f, hy = INVERSE ? 0.5f : -0.5f;
    if (t == 0) {
        float2* p0 = s + C::Layout::phys(fbase);
        const float2 Lv = plat::lds64(p0);
        const float sc = INVERSE ? 0.5f : 1.0f;
        plat::sts64(p0, make_float2(sc * (Lv.x + Lv.y), sc * (Lv.x - Lv.y)));
    }
    static_for<C::R / 2>([&](auto II) {
        constexpr int i = decltype(II)::value;
        const int k = t + 1 + i * C::T;
        float2* pa = s + C::Layout::phys(fbase + k);
        float2* pb = s + C::Layout::phys(fbase + M - k);
        const float2 Av = plat::lds64(pa), Bv = plat::lds64(pb);
        float2 H1, H2, W;
        H1.x = 0.5f * (Av.x + Bv.x);
        H1.y = 0.5f * (Av.y - Bv.y);
        H2.x = hx * (Av.y + Bv.y);
        H2.y = hy * (Av.x - Bv.x);
        static_assert(C::TW == TW_MUFU, "the pair-pass form is only used by the reference-contract API (MUFU twiddles)");
        W = tw_mufu<INVERSE, 2 * M>(k);
        const float2 WH = cmul(W, H2);
        plat::sts64(pa, make_float2(H1.x + WH.x, H1.y + WH.y));
        if (k != M - k) plat::sts64(pb, make_float2(H1.x - WH.x, -H1.y + WH.y));
    });
}

}  // namespace detail
}  // namespace smfft

#include "block_fft_dual.cuh"
