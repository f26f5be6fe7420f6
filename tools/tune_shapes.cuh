// tune_shapes.cuh -- the shape list swept by tools/tune (measurement tool, not product)
#pragma once
#include <initializer_list>
#include <vector>

#include "registry.hpp"

using namespace smfft;
using namespace smfft::host;
using namespace smfft::kernels;

struct Variant {
    KernelEntry k;
    int b, tile_e;
    int hint = 0;        // TileArgs::l2_hint
    int out_off = 0;     // extra byte offset of the output buffer (DRAM bank-aliasing experiment)
    int promo = 3;       // tensor-map L2 promotion: 0 none, 1 64 B, 2 128 B, 3 256 B
    int swz = 1;         // tensor-map swizzle: 1 = 128B (product), 0 = none (staging-only experiments)
    int per_sm = 0;      // CTAs per SM to launch (0 = occupancy limit)
};

extern std::vector<Variant> g_variants;

template <int E, int B, int TILE_E, int STAGES, int MINB, int IO, int REPS = 1, int PF = (IO == IO_TMA ? -1 : 0)>
static void add_one(int hint = 0, int out_off = 0)
{
    if constexpr (TILE_E >= E && E > B && (TILE_E - B) <= 10 && (TILE_E - B) >= 5 && !(IO == IO_TMA_STG && STAGES < 2)) {
        g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 1, IO, TW_LUT, REPS, PF>(), B, TILE_E, hint, out_off});
        if (REPS == 1 && hint == 0 && out_off == 0)
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 0, IO, TW_LUT, REPS, PF>(), B, TILE_E, 0, 0});
    }
}

// same shape without the small-N row skew (E <= 7 only)
template <int E, int B, int TILE_E, int STAGES, int MINB, int IO>
static void add_noskew(int per)
{
    if constexpr (E <= 7 && TILE_E >= E && !(IO == IO_TMA_STG && STAGES < 2)) {
        g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 1, IO, TW_LUT, 1, (IO == IO_TMA ? -1 : 0), false>(), B, TILE_E});
        g_variants.back().per_sm = per;
        g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 0, IO, TW_LUT, 1, (IO == IO_TMA ? -1 : 0), false>(), B, TILE_E});
        g_variants.back().per_sm = per;
    }
}

// late-prefetch variants: more CTAs per SM at the same average load concurrency
template <int E, int B, int TILE_E, int STAGES, int MINB, int PF>
static void add_late(std::initializer_list<int> per_sms)
{
    for (int per : per_sms) {
        const size_t before = g_variants.size();
        add_one<E, B, TILE_E, STAGES, MINB, IO_TMA, 1, PF>();
        add_one<E, B, TILE_E, STAGES, MINB, IO_TMA_STG, 1, PF>();
        for (size_t i = before; i < g_variants.size(); i++) g_variants[i].per_sm = per;
    }
}

// one shape, launched with an explicit number of CTAs per SM (grid = SMs x per_sm): what matters
// to the memory system is the requested load concurrency  per_sm x (STAGES-1) x tile bytes
template <int E, int B, int TILE_E, int STAGES, int MINB>
static void add_shape(std::initializer_list<int> per_sms)
{
    for (int per : per_sms) {
        const size_t before = g_variants.size();
        add_one<E, B, TILE_E, STAGES, MINB, IO_TMA>();
        add_one<E, B, TILE_E, STAGES, MINB, IO_TMA_STG>();
        for (size_t i = before; i < g_variants.size(); i++) g_variants[i].per_sm = per;
    }
}

// R2C / C2R shapes (complex core of 2^E points = real length 2^(E+1)), launched with `per` CTAs per SM
template <int E, int B, int TILE_E, int STAGES, int MINB, int IO, int PF>
static void add_real(std::initializer_list<int> per_sms)
{
    if constexpr (TILE_E >= E && E > B && (TILE_E - B) <= 10 && (TILE_E - B) >= 5) {
        for (int per : per_sms) {
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_R2C, 0, 1, IO, TW_LUT, 1, PF>(), B, TILE_E});
            g_variants.back().per_sm = per;
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2R, 1, 1, IO, TW_LUT, 1, PF>(), B, TILE_E});
            g_variants.back().per_sm = per;
        }
    }
}

// register-direct input (IO_REG): natural-order C2C and R2C; per < 0 = one CTA per tile (non-persistent grid)
template <int E, int B, int TILE_E, int MINB, int PF>
static void add_reg(std::initializer_list<int> per_sms)
{
    if constexpr (TILE_E >= E && E > B && (TILE_E - B) <= 10 && (TILE_E - B) >= 5) {
        for (int per : per_sms) {
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, 1, MINB, MODE_C2C, 0, 1, IO_REG, TW_LUT, 1, PF>(), B, TILE_E});
            g_variants.back().per_sm = per;
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, 1, MINB, MODE_R2C, 0, 1, IO_REG, TW_LUT, 1, PF>(), B, TILE_E});
            g_variants.back().per_sm = per;
        }
    }
}

template <int E>
static void add_reg_size()
{
    add_shape<E, 4, 12, 2, 2>({2});                 // reference outputs for the checks (C2C)
    add_real<E, 4, 12, 2, 2, IO_TMA, -1>({2});      // (R2C / C2R)
    add_reg<E, 4, 10, 8, -1>({8, -1});
    add_reg<E, 4, 10, 8, 0>({4, 6, 8});
    add_reg<E, 4, 11, 4, -1>({4, -1});
    add_reg<E, 4, 11, 4, 0>({2, 3, 4});
    add_reg<E, 4, 11, 6, 0>({6});
    add_reg<E, 4, 12, 2, 0>({1, 2});
    add_reg<E, 4, 12, 3, 0>({3});
    add_reg<E, 4, 12, 2, -1>({2, -1});
    if constexpr (E >= 9) {
        add_reg<E, 5, 11, 4, -1>({4, -1});
        add_reg<E, 5, 11, 4, 0>({2, 3, 4});
        add_reg<E, 5, 11, 6, -1>({6, -1});
        add_reg<E, 5, 12, 2, 0>({1, 2});
        add_reg<E, 5, 12, 3, -1>({3, -1});
        add_reg<E, 5, 12, 4, -1>({4, -1});
    }
}

template <int E>
static void add_real_size()
{
    add_real<E, 4, 12, 2, 2, IO_TMA, -1>({2});  // first: reference output for the checks
    add_real<E, 4, 12, 2, 2, IO_TMA_STG, 0>({2});
    add_real<E, 4, 12, 2, 3, IO_TMA, 1>({3});
    add_real<E, 4, 12, 2, 3, IO_TMA_STG, 1>({3});
    add_real<E, 4, 11, 2, 6, IO_TMA, 1>({4, 5, 6});
    add_real<E, 4, 11, 2, 6, IO_TMA_STG, 1>({4, 5, 6});
    add_real<E, 4, 11, 2, 4, IO_TMA_STG, 0>({3, 4});
    add_real<E, 4, 10, 2, 8, IO_TMA, 1>({6, 8});
    add_real<E, 4, 10, 2, 8, IO_TMA_STG, 1>({6, 8});
    add_real<E, 4, 10, 2, 8, IO_TMA_STG, 0>({6, 8});
    if constexpr (E >= 9 && E <= 11) {
        add_real<E, 5, 12, 2, 2, IO_TMA, 1>({2});
        add_real<E, 5, 12, 2, 2, IO_TMA_STG, 1>({2});
        add_real<E, 5, 12, 2, 3, IO_TMA, 1>({3});
        add_real<E, 5, 12, 2, 3, IO_TMA_STG, 1>({3});
        add_real<E, 5, 11, 2, 4, IO_TMA, 1>({3, 4});
        add_real<E, 5, 11, 2, 4, IO_TMA_STG, 1>({3, 4});
        add_real<E, 5, 11, 2, 6, IO_TMA_STG, 1>({5, 6});
        add_real<E, 5, 11, 2, 4, IO_TMA_STG, 0>({3, 4});
    }
    if constexpr (E == 12) {
        add_real<E, 5, 12, 2, 3, IO_TMA, 1>({3});
        add_real<E, 5, 12, 2, 3, IO_TMA_STG, 1>({3});
        add_real<E, 5, 12, 2, 2, IO_TMA, 1>({2});
    }
}

template <int E>
static void add_size()
{
    add_shape<E, 4, 12, 2, 2>({2});        // first: reference output for the checks
    if constexpr (E <= 8) {
        // product shapes of the small sizes vs the many-small-CTAs + late-prefetch shapes, sustained
        if constexpr (E == 5) add_shape<E, 4, 12, 3, 2>({1});
        if constexpr (E == 6) add_shape<E, 4, 11, 3, 4>({2});
        if constexpr (E == 7) add_shape<E, 4, 10, 2, 8>({6});
        add_late<E, 4, 12, 2, 3, 1>({3});
        add_late<E, 4, 11, 2, 6, 1>({4, 5, 6});
        add_late<E, 4, 10, 2, 8, 1>({6, 8});
        add_shape<E, 4, 11, 2, 4>({3, 4});
    }
    if constexpr (E >= 9) {
        // product R = 16 shapes
        if constexpr (E <= 10) add_late<E, 4, 10, 2, 8, 1>({8});
        if constexpr (E == 11) add_late<E, 4, 11, 2, 6, 1>({5, 6});
        if constexpr (E == 12) add_late<E, 4, 12, 2, 3, 1>({3});
        // R = 32: tile size x CTAs per SM
        add_late<E, 5, 12, 2, 2, 1>({2});
        add_late<E, 5, 12, 2, 3, 1>({3});
        add_late<E, 5, 11, 2, 4, 1>({3, 4});
        add_late<E, 5, 11, 2, 6, 1>({5, 6});
        add_late<E, 5, 10, 2, 8, 1>({6, 8});
        add_late<E, 5, 11, 2, 4, 0>({4});
    }
}

// ---- dual-lane shapes (block_fft_dual.cuh): two transforms per thread, threads = 2^(TILE_E - B - 1) ----
template <int E, int B, int TILE_E, int STAGES, int MINB, int IO, int PF>
static void add_dual_c2c(std::initializer_list<int> per_sms)
{
    if constexpr (TILE_E > E && E - B >= 4 && (TILE_E - B) >= 6 && (TILE_E - B) <= 11) {
        for (int per : per_sms) {
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 1, IO, TW_LUT, 1, PF, true, 1>(), B, TILE_E});
            g_variants.back().per_sm = per;
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 0, IO, TW_LUT, 1, PF, true, 1>(), B, TILE_E});
            g_variants.back().per_sm = per;
        }
    }
}
template <int E, int B, int TILE_E, int STAGES, int MINB, int IO, int PF>
static void add_dual_real(std::initializer_list<int> per_sms)
{
    if constexpr (TILE_E > E && E - B >= 4 && (TILE_E - B) >= 6 && (TILE_E - B) <= 11) {
        for (int per : per_sms) {
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_R2C, 0, 1, IO, TW_LUT, 1, PF, true, 1>(), B, TILE_E});
            g_variants.back().per_sm = per;
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2R, 1, 1, IO, TW_LUT, 1, PF, true, 1>(), B, TILE_E});
            g_variants.back().per_sm = per;
        }
    }
}
// FFT_multiple (100 in-place repetitions, thread-staged): C2C natural order and R2C
template <int E, int B, int TILE_E, int MINB, int DUAL>
static void add_multiple()
{
    if constexpr (TILE_E >= E + DUAL && E - B >= (DUAL ? 4 : 1) && (TILE_E - B - DUAL) >= 5 && (TILE_E - B - DUAL) <= 10) {
        g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, 1, MINB, MODE_C2C, 0, 1, IO_LDG, TW_LUT, 100, 0, true, DUAL>(), B, TILE_E});
        g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, 1, MINB, MODE_R2C, 0, 1, IO_LDG, TW_LUT, 100, 0, true, DUAL>(), B, TILE_E});
    }
}

// the product shape of each kind next to the dual-lane candidates (same box, same run)
template <int E>
static void add_dual_size()
{
    add_shape<E, 4, 12, 2, 2>({2});                 // reference outputs for the checks (C2C)
    add_real<E, 4, 12, 2, 2, IO_TMA, -1>({2});      // (R2C / C2R)
    // product shapes (tuning.hpp)
    if constexpr (E == 8) { add_late<E, 4, 10, 2, 8, 1>({6}); add_real<E, 4, 11, 2, 6, IO_TMA, 1>({4}); }
    if constexpr (E == 9 || E == 10) { add_late<E, 5, 12, 2, 2, 1>({2}); add_real<E, 5, 11, 2, 4, IO_TMA, 1>({4}); add_real<E, 5, 11, 2, 4, IO_TMA_STG, 1>({4}); }
    if constexpr (E == 11) { add_late<E, 4, 11, 2, 6, 1>({5}); add_real<E, 4, 11, 2, 6, IO_TMA, 1>({6}); }
    if constexpr (E == 12) { add_late<E, 5, 12, 2, 3, 1>({3}); add_real<E, 4, 12, 2, 3, IO_TMA, 1>({3}); }
    // dual-lane candidates
    add_dual_c2c<E, 4, 12, 2, 3, IO_TMA, 1>({2, 3});
    add_dual_c2c<E, 4, 12, 2, 3, IO_TMA_STG, 1>({2, 3});
    add_dual_real<E, 4, 12, 2, 3, IO_TMA, 1>({2, 3});
    add_dual_real<E, 4, 12, 2, 3, IO_TMA_STG, 1>({2, 3});
    if constexpr (E <= 10) {
        add_dual_c2c<E, 4, 11, 2, 6, IO_TMA, 1>({4, 6});
        add_dual_real<E, 4, 11, 2, 6, IO_TMA, 1>({4, 6});
        add_dual_real<E, 4, 11, 2, 6, IO_TMA_STG, 1>({4, 6});
    }
    if constexpr (E >= 9 && E <= 11) {
        add_dual_c2c<E, 5, 12, 2, 4, IO_TMA, 1>({3, 4});
        add_dual_real<E, 5, 12, 2, 4, IO_TMA, 1>({3, 4});
        add_dual_real<E, 5, 12, 2, 4, IO_TMA_STG, 1>({3, 4});
    }
    if constexpr (E == 12) {
        add_dual_c2c<E, 4, 13, 2, 1, IO_TMA, 1>({1});
        add_dual_c2c<E, 4, 13, 3, 1, IO_TMA, 2>({1});
        add_dual_real<E, 4, 13, 2, 1, IO_TMA, 1>({1});
        add_dual_real<E, 4, 13, 3, 1, IO_TMA, 2>({1});
        add_dual_c2c<E, 5, 13, 3, 1, IO_TMA, 2>({1});
        add_dual_real<E, 5, 13, 3, 1, IO_TMA, 2>({1});
    }
    // FFT_multiple: product shape vs dual
    using Tm = typename ShapeFor<E, MODE_C2C, 1, 100>::type;
    add_multiple<E, Tm::B, Tm::TILE_E, Tm::MINB, 0>();
    add_multiple<E, 4, 12, 3, 1>();
    add_multiple<E, 4, 11, 6, 1>();
    add_multiple<E, 5, 12, 4, 1>();
    add_multiple<E, 4, 13, 1, 1>();
}
