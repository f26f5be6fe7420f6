// emu_run_cfg.hpp -- one kernel configuration on the SIMT emulator (shared by the emulator translation units).  TESTS ONLY.
#pragma once
#include "emu_runtime.hpp"

#include "smfft/detail/compat_core.cuh"
#include "kernels.cuh"
#include "tuning.hpp"

using namespace smfft;

inline const float2* twiddle_table()
{
    static std::vector<float2> g_tw;
    if (g_tw.empty()) {
        g_tw.resize(kTwiddleTableSize);
        for (int j = 0; j < kTwiddleTableSize; j++) {
            const double a = -2.0 * M_PI * (double)j / (double)kTwiddleTableSize;
            g_tw[j] = make_float2((float)cos(a), (float)sin(a));
        }
    }
    return g_tw.data();
}

template <int E, int B, int F, int MODE, int DIR, int REORDER, int IO, int TW, int STAGES, int REPS, int PF = (IO == kernels::IO_TMA ? -1 : 0), int DUAL = 0>
static int run_cfg(const float2* in, float2* out, long long n_ffts, int grid, double* bank_factor)
{
    using XL = typename std::conditional<B == 5, detail::LayoutSW256, detail::LayoutSW128>::type;
    using C = detail::BlockCfg<E, B, F, DIR, REORDER, TW, detail::LayoutSW128, XL, true, true, DUAL>;
    constexpr int ST = kernels::io_uses_tma(IO) ? STAGES : 1;
    kernels::TileArgs args;
    const long long n_points = n_ffts * C::N;
    const long long n_tiles = (n_points + C::L - 1) / C::L;
    constexpr int BOX = C::L / 16 > 256 ? 256 : C::L / 16;  // as launch.cu: a TMA box holds at most 256 rows
    args.in_map = emu::TensorMapEmu{(unsigned char*)in, n_points / 16, BOX};
    args.out_map = emu::TensorMapEmu{(unsigned char*)out, n_points / 16, BOX};
    args.gin = in;
    args.gout = out;
    args.n_tiles = n_tiles;
    args.n_points = n_points;
    args.tw = twiddle_table();
    args.l2_hint = 0;
    emu::BankStats st;
    if (grid <= 0 || grid > n_tiles) grid = (int)n_tiles;
    emu::launch(grid, C::THREADS, kernels::smem_bytes<C, IO, ST, MODE>(),
                [&](unsigned char* smem) { kernels::tile_kernel_body<C, MODE, IO, ST, REPS, PF>(args, smem); },
                bank_factor ? &st : nullptr);
    if (bank_factor) *bank_factor = st.factor();
    return 0;
}


// dispatch over (dir, reorder, io, tw) for a fixed shape
template <int E, int B, int F, int MODE, int STAGES, int REPS, int PFT = -1, int ARITH = 0>
inline int run_shape(const float2* in, float2* out, long long n_ffts, int dir, int reorder, int io, int tw, int grid,
                     double* bf)
{
#define CASE(D, RO, I, T)                                                     \
    if (dir == D && reorder == RO && io == I && tw == T)                      \
        return run_cfg<E, B, F, MODE, D, RO, I, T, STAGES, REPS, (I == kernels::IO_TMA ? PFT : (PFT < 0 ? 0 : PFT)), ARITH>(in, out, n_ffts, grid, bf);
    if constexpr (MODE == kernels::MODE_C2C) {
        CASE(0, 1, 0, 0) CASE(0, 0, 0, 0) CASE(1, 1, 0, 0) CASE(1, 0, 0, 0)
        CASE(0, 1, 1, 0) CASE(0, 0, 1, 0) CASE(1, 1, 1, 0) CASE(1, 0, 1, 0)
        CASE(0, 1, 0, 1) CASE(0, 0, 0, 1) CASE(1, 1, 0, 1) CASE(1, 0, 0, 1)
        CASE(0, 1, 1, 1) CASE(1, 0, 1, 1)
        if constexpr (STAGES >= 2) { CASE(0, 1, 2, 0) CASE(0, 0, 2, 0) CASE(1, 1, 2, 0) CASE(1, 0, 2, 0) CASE(0, 1, 2, 1) }
    } else if constexpr (MODE == kernels::MODE_R2C) {
        CASE(0, 1, 0, 0) CASE(0, 1, 1, 0) CASE(0, 1, 0, 1)
        if constexpr (STAGES >= 2) { CASE(0, 1, 2, 0) }
    } else {
        CASE(1, 1, 0, 0) CASE(1, 1, 1, 0) CASE(1, 1, 0, 1)
        if constexpr (STAGES >= 2) { CASE(1, 1, 2, 0) }
    }
#undef CASE
    return -1;
}

// the engines behind include/smfft/compat.cuh (reference thread contract: 4 points per thread, linear tile, MUFU twiddles):
// C2C goes through compat::ct_dit -- the very dispatch do_SMFFT_CT_DIT<P> calls -- inside a wrapper with the reference's
// load / sync / FFT / sync / store skeleton; R2C / C2R through the block FFT configuration the Stockham wrappers use
// tests can push the tile off its 16-byte alignment (README.md:12 of the reference asks for none): one float2
inline int& compat_tile_misalign()
{
    static int m = 0;
    return m;
}

template <int E, int DIR, int REORDER>
inline int run_compat_ct(const float2* in, float2* out, long long n_ffts, int reps, double* bank_factor, long long* shuffles)
{
    constexpr int F = E < 7 ? (128 >> E) : 1, L = F << E, T = L / 4;
    if (n_ffts % F) return -2;
    emu::BankStats st;
    const long long n_tiles = n_ffts / F;
    emu::launch((int)n_tiles, T, (size_t)L * 8 + 64,
                [&](unsigned char* smem) {
                    float2* s = reinterpret_cast<float2*>(smem) + compat_tile_misalign();
                    const int tid = plat::tid();
                    const long long base = (long long)plat::bid() * L;
                    if (reps == 0) {  // the body of the external wrapper kernel: tile in global memory
                        compat::ct_dit_external<E, F, DIR, REORDER>(s, in + base, out + base);
                        if (shuffles && tid == 0 && plat::bid() == 0) *shuffles = emu::g_blk->shuffles;
                        return;
                    }
                    for (int q = 0; q < 4; q++) plat::sts64(s + tid + q * T, in[base + tid + q * T]);
                    plat::sync_block();
                    for (int r = 0; r < reps; r++) {
                        compat::ct_dit<E, F, DIR, REORDER>(s);
                        plat::sync_block();
                    }
                    for (int q = 0; q < 4; q++) out[base + tid + q * T] = plat::lds64(s + tid + q * T);
                    if (shuffles && tid == 0 && plat::bid() == 0) *shuffles = emu::g_blk->shuffles;
                },
                bank_factor ? &st : nullptr);
    if (bank_factor) *bank_factor = st.factor();
    return 0;
}

template <int E, int MODE, int DIR, int REORDER>
inline int run_compat(const float2* in, float2* out, long long n_ffts, double* bank_factor)
{
    if constexpr (MODE == kernels::MODE_C2C) {
        return run_compat_ct<E, DIR, REORDER>(in, out, n_ffts, 1, bank_factor, nullptr);
    } else {
    using C = detail::BlockCfg<E, 2, (E < 7 ? (128 >> E) : 1), DIR, REORDER, TW_MUFU, detail::LayoutLinear, detail::LayoutSW128, false>;
    kernels::TileArgs args;
    const long long n_points = n_ffts * C::N;
    args.gin = in;
    args.gout = out;
    args.n_tiles = (n_points + C::L - 1) / C::L;
    args.n_points = n_points;
    args.tw = nullptr;
    args.l2_hint = 0;
    emu::BankStats st;
    emu::launch((int)args.n_tiles, C::THREADS, kernels::smem_bytes<C, kernels::IO_LDG, 1, MODE>(),
                [&](unsigned char* smem) { kernels::tile_kernel_body<C, MODE, kernels::IO_LDG, 1, 1>(args, smem); },
                bank_factor ? &st : nullptr);
    if (bank_factor) *bank_factor = st.factor();
    return 0;
    }
}
