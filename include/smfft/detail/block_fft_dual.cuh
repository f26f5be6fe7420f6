// smfft/detail/block_fft_dual.cuh -- the block FFT with TWO transforms per thread in the packed f32x2 lanes.
//
// Same algorithm, plan and twiddles as block_fft.cuh (Stockham autosort, register-resident radix-R passes; replaces
// do_SMFFT_CT_DIT CT/FFT-GPU-32bit.cu:334-532, do_FFT_Stockham_C2C RC/FFT-GPU-32bit-Stockham.cu:106-266 and
// do_FFT_Stockham_R2C_C2R RC:269-344), but a thread owns point x = t + m*T of transform 2p (lane 0) AND of
// transform 2p+1 (lane 1) of the tile as one cpair (complex.cuh).  Both lanes run the identical index, twiddle and
// butterfly sequence, so
//   * every complex add / multiply is a packed FADD2 / FMUL2 / FFMA2 that serves both transforms: half the
//     floating-point issue slots per point (the large transforms and the real passes are issue-bound, not
//     FMA-pipe-bound: profiles/r01_ncu_summary_z.md, profiles/r01_ab_packed_addsub_z.json);
//   * the exchanges between passes move one 16-byte chunk (re0, re1, im0, im1) per point pair: half the LDS/STS
//     instructions and half the address arithmetic per point, the same bytes;
//   * only the first read and the final write touch the interleaved float2 tile (what TMA loads and stores), with
//     one register move per point to change between (re, im) and (lane 0, lane 1) pairing.
// Exchange layout: the 2N float2 slots of the transform pair are reused as N chunks; chunk x lives at
// x ^ ((x >> B) & 7) -- the SW128 / SW256 idea at chunk granularity: eight consecutive lanes always hit eight
// different 16-byte bank groups, for the column reads x = t + m*T and for every autosort scatter.
#pragma once
#include "block_fft.cuh"

namespace smfft {
namespace detail {

template <class C>
SMFFT_DEV int dual_phys(int x)
{
    return x ^ ((x >> C::B) & 7);
}
template <class C>
SMFFT_DEV cpair dual_lds(const float2* sp, int x)
{
    const float4 q = plat::lds128(sp + 2 * dual_phys<C>(x));
    return cpair{make_float2(q.x, q.y), make_float2(q.z, q.w)};
}
template <class C>
SMFFT_DEV void dual_sts(float2* sp, int x, cpair v)
{
    plat::sts128(sp + 2 * dual_phys<C>(x), make_float4(v.re.x, v.re.y, v.im.x, v.im.y));
}
SMFFT_DEV cpair cpair_from_points(float2 a, float2 b) { return cpair{make_float2(a.x, b.x), make_float2(a.y, b.y)}; }
SMFFT_DEV float2 lane0(cpair v) { return make_float2(v.re.x, v.im.x); }
SMFFT_DEV float2 lane1(cpair v) { return make_float2(v.re.y, v.im.y); }

// ---- interleaved tile <-> registers (sp = first slot of the transform pair, a multiple of 2N >= 256 slots) ----

// tile offset of point x of lane 0; lane 1 is N slots further (N % 128 == 0 keeps the swizzle key)
template <class C>
SMFFT_DEV void dual_load_natural(cpair (&v)[C::R], const float2* sp, int t)
{
    if constexpr (C::T % 128 == 0) {
        const int p0 = LayoutSW128::phys(t);
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = cpair_from_points(plat::lds64(sp + p0 + m * C::T), plat::lds64(sp + p0 + m * C::T + C::N));
        });
    } else {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            const int p = LayoutSW128::phys(t + m * C::T);
            v[m] = cpair_from_points(plat::lds64(sp + p), plat::lds64(sp + p + C::N));
        });
    }
}

template <class C>
SMFFT_DEV void dual_store_natural(const cpair (&v)[C::R], float2* sp, int t)
{
    if constexpr (C::T % 128 == 0) {
        const int p0 = LayoutSW128::phys(t);
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            plat::sts64(sp + p0 + m * C::T, lane0(v[m]));
            plat::sts64(sp + p0 + m * C::T + C::N, lane1(v[m]));
        });
    } else {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            const int p = LayoutSW128::phys(t + m * C::T);
            plat::sts64(sp + p, lane0(v[m]));
            plat::sts64(sp + p + C::N, lane1(v[m]));
        });
    }
}

// fft_reorder = 0: one contiguous run of R points per lane in bit-reversed register order (load_rows_brev)
template <class C>
SMFFT_DEV void dual_load_rows_brev(cpair (&v)[C::R], const float2* sp, int j)
{
    const int row0 = (int)(plat::brev32((unsigned)j) >> (32 - C::A)) * C::R;
    static_for<C::R / 2>([&](auto CI) {
        constexpr int c = decltype(CI)::value;
        const int p = LayoutSW128::phys(row0 + 2 * c);
        const float4 qa = plat::lds128(sp + p), qb = plat::lds128(sp + p + C::N);
        v[brev_c(2 * c, C::B)] = cpair{make_float2(qa.x, qb.x), make_float2(qa.y, qb.y)};
        v[brev_c(2 * c + 1, C::B)] = cpair{make_float2(qa.z, qb.z), make_float2(qa.w, qb.w)};
    });
}

// ---- passes ------------------------------------------------------------------------------------------

template <class C, int PIDX>
SMFFT_DEV void dual_pass_compute(cpair (&v)[C::R], int vt, const float2* tw)
{
    constexpr int c = C::radix_log2(PIDX), r = 1 << c, U = C::R / r;
    constexpr int NS = 1 << C::ns_log2(PIDX);
    if constexpr (NS > 1) {
        constexpr int WN = NS * r;
        cpair pw[r];
        make_twiddle_powers<C::DIR, C::TW, WN, r>(pw, vt & (NS - 1), tw + C::tw_offset(PIDX));
        constexpr int D = NS > C::T ? NS / C::T : 1;  // see fft_pass_compute
        static_for<U>([&](auto UI) {
            constexpr int u = decltype(UI)::value;
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                if constexpr (q >= 1) {
                    cpair x = cmul(v[u + q * U], pw[q]);
                    if constexpr (D > 1) x = mul_wconst<C::DIR, ((u % D) * q) % (D * r), D * r>(x);
                    v[u + q * U] = x;
                }
            });
        });
    }
    static_for<U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        dft_regs<0, C::DIR, r, u, U, C::R>(v);
    });
}

template <class C, int PIDX>
SMFFT_DEV void dual_pass_scatter(const cpair (&v)[C::R], float2* sp, int vt)
{
    constexpr int c = C::radix_log2(PIDX), r = 1 << c, U = C::R / r;
    constexpr int LNS = C::ns_log2(PIDX), NS = 1 << LNS;
    static_for<U>([&](auto UI) {
        constexpr int u = decltype(UI)::value;
        const int j = vt + u * C::T;
        const int xb = ((j >> LNS) << (LNS + c)) + (j & (NS - 1));
        if constexpr (NS >= (8 << C::B)) {
            const int p0 = dual_phys<C>(xb);  // q*NS leaves the key bits B..B+2 and the chunk bits 0..2 alone
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                const cpair x = v[u + q * U];
                plat::sts128(sp + 2 * (p0 + q * NS), make_float4(x.re.x, x.re.y, x.im.x, x.im.y));
            });
        } else {
            static_for<r>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                dual_sts<C>(sp, xb + q * NS, v[u + q * U]);
            });
        }
    });
}

template <class C>
SMFFT_DEV void dual_load_exchange(cpair (&v)[C::R], const float2* sp, int t)
{
    if constexpr (C::T % (8 << C::B) == 0) {
        const int p0 = dual_phys<C>(t);
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            const float4 q = plat::lds128(sp + 2 * (p0 + m * C::T));
            v[m] = cpair{make_float2(q.x, q.y), make_float2(q.z, q.w)};
        });
    } else {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            v[m] = dual_lds<C>(sp, t + m * C::T);
        });
    }
}

template <class C, int PIDX, class Hook>
SMFFT_DEV void dual_run_passes(cpair (&v)[C::R], float2* sp, int vt, int t, const float2* tw, Hook&& hook)
{
    dual_pass_compute<C, PIDX>(v, vt, tw);
    if constexpr (PIDX + 1 < C::P) {
        plat::sync_block();  // every thread has finished reading the previous state of the tile
        if constexpr (PIDX == (std::remove_reference<Hook>::type::PASS < C::P - 2 ? std::remove_reference<Hook>::type::PASS : C::P - 2)) hook();
        dual_pass_scatter<C, PIDX>(v, sp, vt);
        plat::sync_block();
        dual_load_exchange<C>(v, sp, t);
        dual_run_passes<C, PIDX + 1>(v, sp, t, t, tw, hook);
    }
}

// ---- R2C / C2R (formulas and reference lines: real_pass_regs / r2c_tail_regs in block_fft.cuh) ---------------

// W_{2N}^t / 2 of the real pass, both lanes
template <class C, int INVERSE>
SMFFT_DEV cpair dual_real_twiddle(const float2* tw, int t)
{
    float2 wt;
    if constexpr (C::TW == TW_LUT) {
        wt = plat::lds64(tw + C::TW_C2C_ENTRIES + t);
    } else {
        wt = tw_mufu<INVERSE, 2 * C::N>(t);
        wt.x *= 0.5f;
        wt.y *= 0.5f;
    }
    return cdup(wt);
}

// C2R head: v[m] = Y[t + m*T] (both lanes) -> Z[t + m*T]; partners Y[N - k] come from the read-only interleaved tile
template <class C>
SMFFT_DEV void dual_c2r_head(cpair (&v)[C::R], const float2* sp, int t, const float2* tw)
{
    static_assert(2 * C::R <= 64, "constant twiddles W_{2R}^m come from the W_64 table");
    const cpair wt = dual_real_twiddle<C, 1>(tw, t);
    cpair wm[C::R];
    static_for<C::R / 2>([&](auto MI) {
        constexpr int m = decltype(MI)::value;
        wm[m] = mul_wconst<1, m, 2 * C::R>(wt);
        wm[m + C::R / 2] = mul_wconst<1, 1, 4>(wm[m]);
    });
    const int xp = C::N - t;  // partner of k = t + m*T is (N - t) - m*T
    static_for<C::R>([&](auto MI) {
        constexpr int m = decltype(MI)::value;
        const int x = (m == 0 && t == 0) ? 0 : xp - m * C::T;
        const int p = LayoutSW128::phys(x);
        const cpair Bv = cpair_from_points(plat::lds64(sp + p), plat::lds64(sp + p + C::N));
        const cpair A = v[m], Wh = wm[m];
        const float2 sx = pk::add(A.re, Bv.re), sy = pk::add(A.im, Bv.im), dx = pk::sub(A.re, Bv.re), dy = pk::sub(A.im, Bv.im);
        // inverse: Z = ( sx/2 - (Wh.x sy + Wh.y dx),  dy/2 - (Wh.y sy - Wh.x dx) )
        cpair o;
        o.re = pk::fma(pk::neg(Wh.re), sy, pk::fma(pk::neg(Wh.im), dx, pk::mul(sx, pk::bc(0.5f))));
        o.im = pk::fma(Wh.re, dx, pk::fma(pk::neg(Wh.im), sy, pk::mul(dy, pk::bc(0.5f))));
        if constexpr (m == 0) {
            if (t == 0) {  // bin 0 un-packed (RC:280-286)
                o.re = pk::mul(pk::add(A.re, A.im), pk::bc(0.5f));
                o.im = pk::mul(pk::sub(A.re, A.im), pk::bc(0.5f));
            }
        }
        v[m] = o;
    });
}

// R2C tail, pair form: on entry v[m] = Z[t + m*T]; on return v[i] = X[t + i*T], v[R/2 + i] = X[r2c_hi_index(t, i)].
// The upper half of every thread's values goes through the chunk layout (its own slots of the last exchange).
template <class C>
SMFFT_DEV void dual_r2c_tail(cpair (&v)[C::R], float2* sp, int t, const float2* tw)
{
    static_assert(2 * C::R <= 64, "constant twiddles W_{2R}^i come from the W_64 table");
    constexpr int H = C::R / 2;
    static_for<H>([&](auto II) {
        constexpr int m = H + decltype(II)::value;
        dual_sts<C>(sp, t + m * C::T, v[m]);
    });
    const cpair wt = dual_real_twiddle<C, 0>(tw, t);
    const cpair zmid = v[H];
    plat::sync_block();
    static_for<H>([&](auto II) {
        constexpr int i = decltype(II)::value;
        const cpair Wh = mul_wconst<0, i, 2 * C::R>(wt);
        const cpair A = v[i];
        const int xb = (i == 0 && t == 0) ? C::N / 2 : C::N - t - i * C::T;
        const cpair Bv = dual_lds<C>(sp, xb);
        const float2 sx = pk::add(A.re, Bv.re), sy = pk::sub(A.im, Bv.im), dx = pk::sub(A.re, Bv.re), dy = pk::add(A.im, Bv.im);
        const float2 px = pk::fma(Wh.re, dy, pk::mul(Wh.im, dx)), py = pk::fma(Wh.im, dy, pk::mul(pk::neg(Wh.re), dx));
        cpair lo{pk::fma(sx, pk::bc(0.5f), px), pk::fma(sy, pk::bc(0.5f), py)};
        cpair hi{pk::fma(sx, pk::bc(0.5f), pk::neg(px)), pk::fma(sy, pk::bc(-0.5f), py)};
        if constexpr (i == 0) {
            if (t == 0) {
                lo = cpair{pk::add(A.re, A.im), pk::sub(A.re, A.im)};
                hi = cpair{zmid.re, pk::neg(zmid.im)};
            }
        }
        v[i] = lo;
        v[H + i] = hi;
    });
}

// load (+ C2R head) + all passes; result in registers: v[m] = X[t + m*T] of transforms 2p (lane 0) and 2p+1 (lane 1)
template <class C, int XF, class Hook>
SMFFT_DEV void dual_fft_regs(cpair (&v)[C::R], float2* sp, int t, const float2* tw, Hook&& hook)
{
    int vt = t;
    if constexpr (C::REORDER) {
        dual_load_natural<C>(v, sp, t);
        if constexpr (XF == XF_C2R) dual_c2r_head<C>(v, sp, t, tw);  // tile is read-only here: no barrier
    } else {
        vt = noreorder_vid<C>(t);
        dual_load_rows_brev<C>(v, sp, vt);
    }
    dual_run_passes<C, 0>(v, sp, vt, t, tw, hook);
}

template <class C, int XF, class Hook>
SMFFT_DEV void dual_fft_tile(float2* s, const float2* tw, Hook&& hook)
{
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    float2* sp = s + ((tid >> C::A) << (C::E + 1));
    cpair v[C::R];
    dual_fft_regs<C, XF>(v, sp, t, tw, hook);
    if constexpr (XF == XF_R2C) dual_r2c_tail<C>(v, sp, t, tw);
    plat::sync_block();  // the interleaved result overwrites chunks other threads have just read
    if constexpr (XF == XF_R2C) {
        static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            const int p = LayoutSW128::phys(r2c_out_index<C>(t, m));
            plat::sts64(sp + p, lane0(v[m]));
            plat::sts64(sp + p + C::N, lane1(v[m]));
        });
    } else {
        dual_store_natural<C>(v, sp, t);
    }
}

template <class C, int XF, class Hook>
SMFFT_DEV void dual_fft_tile_to_global(float2* s, const float2* tw, float2* __restrict__ g, long long valid, Hook&& hook)
{
    const int tid = plat::tid();
    const int t = tid & (C::T - 1);
    const int fbase = (tid >> C::A) << (C::E + 1);
    cpair v[C::R];
    dual_fft_regs<C, XF>(v, s + fbase, t, tw, hook);
    if constexpr (XF == XF_R2C) dual_r2c_tail<C>(v, s + fbase, t, tw);
    auto index = [&](auto M) {
        constexpr int m = decltype(M)::value;
        if constexpr (XF == XF_R2C)
            return fbase + r2c_out_index<C>(t, m);
        else
            return fbase + t + m * C::T;
    };
    if (valid >= C::L) {
        static_for<C::R>([&](auto M) {
            const int x = index(M);
            plat::stg64_stream(g + x, lane0(v[decltype(M)::value]));
            plat::stg64_stream(g + x + C::N, lane1(v[decltype(M)::value]));
        });
    } else {
        static_for<C::R>([&](auto M) {
            const int x = index(M);
            if (x < valid) plat::stg64_stream(g + x, lane0(v[decltype(M)::value]));
            if (x + C::N < valid) plat::stg64_stream(g + x + C::N, lane1(v[decltype(M)::value]));
        });
    }
}

}  // namespace detail
}  // namespace smfft
