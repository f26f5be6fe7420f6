// registry.hpp -- table of compiled kernel instances, filled by the per-size translation units.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "tuning.hpp"

namespace smfft {
namespace host {

struct KernelEntry {
    int mode, e, dir, reorder, io, tw, reps;  // lookup key (e = log2 of the complex length)
    int tile_points, threads, smem_bytes, minb, stages;
    int ctas;    // CTAs per SM to launch (0 = occupancy limit)
    int prefer;  // 1 = the default instance of this transform (exactly one per key)
    int tma_best; // 1 = the measured best TMA staging of this transform (= prefer unless a register-direct shape is the default)
    int pf;      // pass after which the next tile is prefetched (-1: top of the iteration)
    int skew;    // small-transform bank de-conflicting on (1) / off (0)
    int dual;    // 1 = two transforms per thread in the packed f32x2 lanes (block_fft_dual.cuh)
    int variant; // 0 = the instance of the static table; 1, 2 = alternates (candidates of the first-use selection; io = 4 / 5 for the register-direct ones)
    const void* func;
};

struct EntryList {
    const KernelEntry* entries;
    int count;
};

// one instance with an explicit shape (used by tools/tune.cu to sweep shapes)
template <int E, int B, int TILE_E, int STAGES, int MINB, int MODE, int DIR, int REORDER, int IO, int TW, int REPS,
          int PF = (IO == kernels::IO_TMA ? -1 : 0), bool SKEW = true, int DUAL = 0>
KernelEntry make_entry_shape()
{
    using XL = typename std::conditional<B == 5, detail::LayoutSW256, detail::LayoutSW128>::type;
    using C = detail::BlockCfg<E, B, (1 << (TILE_E - E)), DIR, REORDER, TW, detail::LayoutSW128, XL, true, SKEW, DUAL>;
    constexpr int ST = kernels::io_uses_tma(IO) ? STAGES : 1;
    KernelEntry k;
    k.mode = MODE; k.e = E; k.dir = DIR; k.reorder = REORDER; k.io = IO; k.tw = TW; k.reps = REPS;
    k.tile_points = C::L; k.threads = C::THREADS; k.smem_bytes = kernels::smem_bytes<C, IO, ST, MODE>();
    k.minb = MINB; k.stages = ST; k.ctas = 0; k.prefer = 0; k.tma_best = 0;
    k.pf = PF; k.skew = SKEW; k.dual = DUAL;  // 0 scalar, 1 dual-lane, 2 packed add / subtract
    k.variant = 0;
    k.func = reinterpret_cast<const void*>(&kernels::smfft_tile_kernel<C, MODE, IO, ST, REPS, MINB, PF>);
    return k;
}

// the product instance: shape from tuning.hpp
template <int E, int MODE, int DIR, int REORDER, int IO, int TW, int REPS>
KernelEntry make_entry()
{
    using Tn = typename kernels::ShapeFor<E, MODE, REORDER, REPS>::type;
    // one buffer and TMA stores: the refill can only follow the store of the same buffer (top of the iteration)
    constexpr int PF = IO == kernels::IO_TMA ? (Tn::STAGES == 1 ? -1 : Tn::PF) : (Tn::PF < 0 ? 0 : Tn::PF);
    constexpr int ARITH = kernels::ArithFor<E, MODE, REORDER, REPS>::value;
    KernelEntry k = make_entry_shape<E, Tn::B, Tn::TILE_E, Tn::STAGES, Tn::MINB, MODE, DIR, REORDER, IO, TW, REPS, PF, true, ARITH>();
    k.ctas = REPS > 1 ? 0 : Tn::CTAS;  // FFT_multiple is compute-bound: fill the SM
    constexpr bool stg = (Tn::STAGES >= 2 || MODE == kernels::MODE_C2C) && (MODE == kernels::MODE_R2C ? Tn::STG_R2C : MODE == kernels::MODE_C2R ? Tn::STG_C2R : Tn::STG);
    k.tma_best = (IO == kernels::IO_TMA_STG) ? stg : (IO == kernels::IO_TMA ? !stg : 0);
    k.prefer = k.tma_best;
    if (kernels::RegDirect<E>::ON && kernels::RegDirect<E>::PREFER && MODE == kernels::MODE_C2C && REORDER == 1 && REPS == 1 && TW == TW_LUT) k.prefer = 0;
    return k;
}

// the register-direct instances (natural-order C2C external): shapes A / B from kernels::RegDirect
template <int E, int DIR, int TW, int WHICH = 0>
KernelEntry make_entry_reg()
{
    using Rd = kernels::RegDirect<E>;
    constexpr int B = WHICH ? Rd::B_B : Rd::B, TILE_E = WHICH ? Rd::TILE_E_B : Rd::TILE_E, MINB = WHICH ? Rd::MINB_B : Rd::MINB;
    KernelEntry k = make_entry_shape<E, B, TILE_E, 1, MINB, kernels::MODE_C2C, DIR, 1, kernels::IO_REG, TW, 1, -1>();
    k.ctas = -1;              // one CTA per tile (non-persistent grid)
    k.prefer = WHICH == 0 && Rd::PREFER && TW == TW_LUT;    // 1 = the default for this size (table twiddles); the TMA instances stay reachable with io = 2 / 3
    k.variant = 1 + WHICH;
    return k;
}

// register-direct R2C (real length 2^(E+1)): shape A of kernels::RegDirect, table twiddles -- an alternate (io = 4, first-use selection)
template <int E>
KernelEntry make_entry_reg_r2c()
{
    using Rd = kernels::RegDirect<E>;
    KernelEntry k = make_entry_shape<E, Rd::B, Rd::TILE_E, 1, Rd::MINB, kernels::MODE_R2C, 0, 1, kernels::IO_REG, TW_LUT, 1, -1>();
    k.ctas = -1;
    k.variant = 1;
    return k;
}

// alternates of the TMA path where single launches and the sustained step disagree (tuning.hpp): candidates of the first-use selection
template <int E, int DIR>
KernelEntry make_entry_alt_tma()
{
    using namespace kernels;
    static_assert(E == 12 || E == 5, "alternates exist for 4096 points (R = 32 plan) and 32 points (one large CTA per SM)");
    KernelEntry k;
    if constexpr (E == 12) {
        using Tn = TuningR32<12>;  // [32,32,4]: wins single launches, loses under the power cap (profiles/r01_bench_sustained_r32_vs_r16_4096.json)
        k = make_entry_shape<12, Tn::B, Tn::TILE_E, Tn::STAGES, Tn::MINB, MODE_C2C, DIR, 1, IO_TMA, TW_LUT, 1, Tn::PF, true, 0>();
        k.ctas = Tn::CTAS;
    } else {
        k = make_entry_shape<5, 4, 12, 3, 1, MODE_C2C, DIR, 1, IO_TMA, TW_LUT, 1, 1, true, 0>();  // one 32 KB-tile CTA per SM, three stages
        k.ctas = 1;
    }
    k.variant = 1;
    return k;
}

// every instance the C ABI can dispatch to for one size
template <int E>
EntryList build_entries()
{
    using namespace kernels;
    // filled exactly once, by whichever thread gets here first (C++11 guarantees the initialisation of a function-local
    // static is thread-safe); read-only afterwards
    struct Table {
        KernelEntry tab[124];
        int n = 0;
    };
    static const Table table = [] {
        Table t;
        KernelEntry* tab = t.tab;
        int i = 0;
#define SMFFT_ADD(...) tab[i++] = make_entry<E, __VA_ARGS__>()
        // C2C external (FFT_external_benchmark): dir x reorder x io x twiddle
        SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_MUFU, 1);
        // register-output staging (TMA in, STG out): C2C always (one-stage shapes refill behind the final exchange), the real
        // transforms where the shape has two tile buffers
        SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA_STG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA_STG, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA_STG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA_STG, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA_STG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA_STG, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA_STG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA_STG, TW_MUFU, 1);
        if constexpr (ShapeFor<E, MODE_C2R, 1, 1>::type::STAGES >= 2) {
            SMFFT_ADD(MODE_C2R, 1, 1, IO_TMA_STG, TW_LUT, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_TMA_STG, TW_MUFU, 1);
        }
        if constexpr (ShapeFor<E, MODE_R2C, 1, 1>::type::STAGES >= 2) {
            SMFFT_ADD(MODE_R2C, 0, 1, IO_TMA_STG, TW_LUT, 1); SMFFT_ADD(MODE_R2C, 0, 1, IO_TMA_STG, TW_MUFU, 1);
        }
        // register-direct input (IO_REG): 1024-point natural-order C2C -- R = 32, one warp per transform, two transforms per
        // 64-thread CTA, one CTA per tile, launched with the driver's default L1 carve-out (tuning.hpp, RegDirect)
        if constexpr (RegDirect<E>::ON) {
            tab[i++] = make_entry_reg<E, 0, TW_LUT>(); tab[i++] = make_entry_reg<E, 1, TW_LUT>();
            tab[i++] = make_entry_reg<E, 0, TW_MUFU>(); tab[i++] = make_entry_reg<E, 1, TW_MUFU>();
        }
        if constexpr (RegDirect<E>::ON) tab[i++] = make_entry_reg_r2c<E>();
        if constexpr (RegDirect<E>::ON_B) {
            tab[i++] = make_entry_reg<E, 0, TW_LUT, 1>(); tab[i++] = make_entry_reg<E, 1, TW_LUT, 1>();
        }
        if constexpr (E == 12 || E == 5) {
            tab[i++] = make_entry_alt_tma<E, 0>(); tab[i++] = make_entry_alt_tma<E, 1>();
        }
        // C2C multiple (FFT_multiple_benchmark, 100 reps in place): compute-bound, LDG staging only
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_LUT, 100); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_LUT, 100);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_LUT, 100); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_LUT, 100);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_MUFU, 100); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_MUFU, 100);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_MUFU, 100); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_MUFU, 100);
        // R2C / C2R external (complex core 2^E, real length 2^(E+1))
        SMFFT_ADD(MODE_R2C, 0, 1, IO_TMA, TW_LUT, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_TMA, TW_LUT, 1);
        SMFFT_ADD(MODE_R2C, 0, 1, IO_TMA, TW_MUFU, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_TMA, TW_MUFU, 1);
        SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_LUT, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_LDG, TW_LUT, 1);
        SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_MUFU, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_LDG, TW_MUFU, 1);
        // R2C multiple (forward only, as RC:445-457)
        SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_LUT, 100); SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_MUFU, 100);
        // the same repeated path with THREE repetitions: F(F(F(x))) stays finite, so its values can be checked against the
        // oracle on the GPU (smfft_exec_repeated; the 100-rep instances overflow by design)
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_LUT, 3); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_LUT, 3);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_LUT, 3); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_LUT, 3);
        SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_LUT, 3);
#undef SMFFT_ADD
        t.n = i;
        return t;
    }();
    return EntryList{table.tab, table.n};
}

// 8192 / 16384 points: C2C external only (table and MUFU twiddles, TMA and thread staging)
template <int E>
EntryList build_entries_large()
{
    using namespace kernels;
    struct Table {
        KernelEntry tab[32];
        int n = 0;
    };
    static const Table table = [] {
        Table t;
        KernelEntry* tab = t.tab;
        int i = 0;
#define SMFFT_ADD(...) tab[i++] = make_entry<E, __VA_ARGS__>()
        // 16384 reals on the 8192-point core: R2C / C2R external, TMA and thread staging
        if constexpr (E == 13) {
            SMFFT_ADD(MODE_R2C, 0, 1, IO_TMA, TW_LUT, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_TMA, TW_LUT, 1);
            SMFFT_ADD(MODE_R2C, 0, 1, IO_TMA, TW_MUFU, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_TMA, TW_MUFU, 1);
            SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_LUT, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_LDG, TW_LUT, 1);
            SMFFT_ADD(MODE_R2C, 0, 1, IO_LDG, TW_MUFU, 1); SMFFT_ADD(MODE_C2R, 1, 1, IO_LDG, TW_MUFU, 1);
        }
        // 16384 points, one 128 KB buffer: TMA in, results out from registers, the refill issued behind the final exchange
        if constexpr (Tuning<E>::STAGES == 1) {
            SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA_STG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA_STG, TW_LUT, 1);
            SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA_STG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA_STG, TW_LUT, 1);
            SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA_STG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA_STG, TW_MUFU, 1);
            SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA_STG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA_STG, TW_MUFU, 1);
        }
        SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_TMA, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_TMA, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_TMA, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_TMA, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_LUT, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_LUT, 1);
        SMFFT_ADD(MODE_C2C, 0, 1, IO_LDG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 0, 0, IO_LDG, TW_MUFU, 1);
        SMFFT_ADD(MODE_C2C, 1, 1, IO_LDG, TW_MUFU, 1); SMFFT_ADD(MODE_C2C, 1, 0, IO_LDG, TW_MUFU, 1);
#undef SMFFT_ADD
        t.n = i;
        return t;
    }();
    return EntryList{table.tab, table.n};
}

// defined one per translation unit (inst_e5.cu ... inst_e14.cu) so the sizes compile in parallel
EntryList entries_e5();
EntryList entries_e6();
EntryList entries_e7();
EntryList entries_e8();
EntryList entries_e9();
EntryList entries_e10();
EntryList entries_e11();
EntryList entries_e12();
EntryList entries_e13();
EntryList entries_e14();

}  // namespace host
}  // namespace smfft
