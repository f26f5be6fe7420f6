// emu_alt.cpp -- non-product shapes on the SIMT emulator: reference-contract configuration, other radix plans, late
// prefetch, register-direct input, reversed plans (own translation unit: compiles in parallel).  TESTS ONLY.
#include "emu_run_cfg.hpp"

extern "C" {

int emu_run_compat(const void* in, void* out, int e, long long n_ffts, int mode, int dir, int reorder, double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
#define CC(E)                                                                                      \
    if (e == E) {                                                                                  \
        if (mode == 0 && dir == 0 && reorder == 1) return run_compat<E, 0, 0, 1>(i, o, n_ffts, bank_factor); \
        if (mode == 0 && dir == 0 && reorder == 0) return run_compat<E, 0, 0, 0>(i, o, n_ffts, bank_factor); \
        if (mode == 0 && dir == 1 && reorder == 1) return run_compat<E, 0, 1, 1>(i, o, n_ffts, bank_factor); \
        if (mode == 0 && dir == 1 && reorder == 0) return run_compat<E, 0, 1, 0>(i, o, n_ffts, bank_factor); \
        if (mode == 1) return run_compat<E, 1, 0, 1>(i, o, n_ffts, bank_factor);                     \
        if (mode == 2) return run_compat<E, 2, 1, 1>(i, o, n_ffts, bank_factor);                     \
    }
    CC(5) CC(6) CC(7) CC(8) CC(9) CC(10) CC(11) CC(12)
#undef CC
    return -1;
}

void emu_set_compat_misalign(int float2s) { compat_tile_misalign() = float2s ? 1 : 0; }

// the reference-contract C2C device function (compat::ct_dit, i.e. do_SMFFT_CT_DIT) with `reps` in-place repetitions;
// also reports the bank-conflict factor of its shared-memory accesses and the warp shuffles of one block
int emu_run_compat_ct(const void* in, void* out, int e, long long n_ffts, int dir, int reorder, int reps, double* bank_factor,
                      long long* shuffles)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
#define CT(E)                                                                                                        \
    if (e == E) {                                                                                                    \
        if (dir == 0 && reorder == 1) return run_compat_ct<E, 0, 1>(i, o, n_ffts, reps, bank_factor, shuffles);      \
        if (dir == 0 && reorder == 0) return run_compat_ct<E, 0, 0>(i, o, n_ffts, reps, bank_factor, shuffles);      \
        if (dir == 1 && reorder == 1) return run_compat_ct<E, 1, 1>(i, o, n_ffts, reps, bank_factor, shuffles);      \
        if (dir == 1 && reorder == 0) return run_compat_ct<E, 1, 0>(i, o, n_ffts, reps, bank_factor, shuffles);      \
    }
    CT(5) CT(6) CT(7) CT(8) CT(9) CT(10) CT(11) CT(12)
#undef CT
    return -1;
}

// alternative shapes: exercise the generic pass machinery (other radices, tile sizes, stage counts)
int emu_run_alt(const void* in, void* out, int variant, long long n_ffts, int dir, int reorder, int io, int tw, int grid,
                double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    switch (variant) {
        case 0: return run_shape<10, 3, 1, 0, 1, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 1024 = 8*8*8*2, R=8
        case 1: return run_shape<10, 5, 4, 0, 3, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 1024 = 32*32, 3 stages
        case 2: return run_shape<9, 3, 4, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);    // 512 = 8*8*8
        case 3: return run_shape<7, 2, 8, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);    // 128, R=4 (compat shape)
        case 4: return run_shape<12, 4, 2, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 2 x 4096 per tile
        case 5: return run_shape<6, 3, 16, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 64 = 8*8
        case 6: return run_shape<5, 2, 16, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 32 = 4*4*2
        case 7: return run_shape<11, 5, 1, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 2048 = 32*32*2
    }
    return -1;
}

// late prefetch (the refill of the previous buffer issued after pass PF instead of at the first barrier)
int emu_run_late(const void* in, void* out, int variant, long long n_ffts, int grid)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    switch (variant) {
        case 0: return run_cfg<12, 4, 1, 0, 0, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);
        case 1: return run_cfg<12, 4, 1, 0, 1, 0, kernels::IO_TMA_STG, TW_LUT, 2, 1, 2>(i, o, n_ffts, grid, nullptr);
        case 2: return run_cfg<10, 4, 4, 0, 0, 0, kernels::IO_TMA, TW_LUT, 2, 1, 2>(i, o, n_ffts, grid, nullptr);
        case 3: return run_cfg<8, 4, 8, 0, 0, 1, kernels::IO_TMA_STG, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);   // P = 2: clamps to pass 0
        case 4: return run_cfg<10, 4, 4, 1, 0, 1, kernels::IO_TMA_STG, TW_LUT, 3, 1, 1>(i, o, n_ffts, grid, nullptr);  // R2C, three buffers
        case 5: return run_cfg<11, 4, 2, 2, 1, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);      // C2R
        // register-direct input (IO_REG): global -> registers, with and without the software prefetch
        case 6: return run_cfg<10, 4, 1, 0, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, 0>(i, o, n_ffts, grid, nullptr);
        case 7: return run_cfg<9, 5, 4, 0, 1, 1, kernels::IO_REG, TW_LUT, 1, 1, -1>(i, o, n_ffts, grid, nullptr);
        case 8: return run_cfg<10, 4, 2, 1, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, 0>(i, o, n_ffts, grid, nullptr);     // R2C
        case 9: return run_cfg<7, 4, 8, 1, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, -1>(i, o, n_ffts, grid, nullptr);     // R2C
        // reversed plan [8,16,16] with the mirrored C2R head (arith flags 6 = packed add/sub + reversed), TMA / LDG / TMA_STG / MUFU
        case 10: return run_cfg<11, 4, 1, 2, 1, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1, 6>(i, o, n_ffts, grid, nullptr);
        case 11: return run_cfg<11, 4, 2, 2, 1, 1, kernels::IO_LDG, TW_LUT, 1, 1, 0, 4>(i, o, n_ffts, grid, nullptr);
        case 12: return run_cfg<11, 4, 1, 2, 1, 1, kernels::IO_TMA_STG, TW_MUFU, 2, 1, 1, 6>(i, o, n_ffts, grid, nullptr);
        // reversed plan, plain C2C (both directions) and R2C through the same machinery
        case 13: return run_cfg<11, 4, 1, 0, 0, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1, 4>(i, o, n_ffts, grid, nullptr);
        case 14: return run_cfg<11, 4, 1, 0, 1, 1, kernels::IO_LDG, TW_LUT, 1, 1, 0, 4>(i, o, n_ffts, grid, nullptr);
        case 15: return run_cfg<7, 4, 8, 0, 0, 1, kernels::IO_TMA, TW_LUT, 2, 1, -1, 4>(i, o, n_ffts, grid, nullptr);   // [8,16]
        // mirrored R2C ownership with more than one butterfly pair per thread: U = 8 (R = 32, last radix 4), U = 8 (R = 16, last radix 2), U = 4
        case 16: return run_cfg<12, 5, 1, 1, 0, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1, 2>(i, o, n_ffts, grid, nullptr);
        case 17: return run_cfg<9, 4, 4, 1, 0, 1, kernels::IO_TMA_STG, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);
        case 18: return run_cfg<10, 4, 2, 1, 0, 1, kernels::IO_LDG, TW_MUFU, 1, 1, 0, 2>(i, o, n_ffts, grid, nullptr);
        // reversed plans with a radix-4 first pass: [4,32,32] C2C and mirrored C2R (U = 8), [4,16,16] mirrored C2R (U = 4)
        case 19: return run_cfg<12, 5, 1, 0, 0, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1, 4>(i, o, n_ffts, grid, nullptr);
        case 20: return run_cfg<12, 5, 1, 2, 1, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1, 6>(i, o, n_ffts, grid, nullptr);
        case 21: return run_cfg<10, 4, 2, 2, 1, 1, kernels::IO_TMA_STG, TW_LUT, 2, 1, 1, 6>(i, o, n_ffts, grid, nullptr);
        case 22: return run_cfg<10, 4, 1, 0, 1, 1, kernels::IO_LDG, TW_MUFU, 1, 1, 0, 4>(i, o, n_ffts, grid, nullptr);
        // 16384 reals: the 8192-point core [32,32,8] (three buffers) with the mirrored real passes -- R2C plain plan, C2R reversed [8,32,32]
        case 23: return run_cfg<13, 5, 1, 1, 0, 1, kernels::IO_TMA, TW_LUT, 3, 1, 1, 2>(i, o, n_ffts, grid, nullptr);
        case 24: return run_cfg<13, 5, 1, 2, 1, 1, kernels::IO_TMA, TW_LUT, 3, 1, 1, 6>(i, o, n_ffts, grid, nullptr);
    }
    return -1;
}

}  // extern "C"

// the product's ALTERNATE instances (tuning.hpp RegDirect shapes A / B, alternate TMA shapes): candidates of the first-use
// selection.  which: 0 = register-direct A, 1 = register-direct B, 2 = alternate TMA shape (4096 points R = 32, 32 points one large CTA)
template <int E, int WHICH>
static int run_reg_alt(const float2* i, float2* o, long long n_ffts, int dir, int grid)
{
    using Rd = kernels::RegDirect<E>;
    constexpr int B = WHICH ? Rd::B_B : Rd::B, TILE_E = WHICH ? Rd::TILE_E_B : Rd::TILE_E;
    if constexpr (WHICH ? Rd::ON_B : Rd::ON) {
        if (dir == 0) return run_cfg<E, B, (1 << (TILE_E - E)), 0, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, -1>(i, o, n_ffts, grid, nullptr);
        return run_cfg<E, B, (1 << (TILE_E - E)), 0, 1, 1, kernels::IO_REG, TW_LUT, 1, 1, -1>(i, o, n_ffts, grid, nullptr);
    }
    return -3;
}

extern "C" {

int emu_run_alternate(const void* in, void* out, int e, int which, long long n_ffts, int dir, int grid)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    if (e == 13) {  // 8192 points (Tuning<13>): which = reorder | io << 1 (io: 0 TMA as two 256-row boxes per tile, 1 thread staging)
        using T13 = kernels::Tuning<13>;
        constexpr int A13 = kernels::ArithFor<13, 0, 1, 1>::value;
        const int reorder = which & 1, io = which >> 1;
#define C13(D, RO, IO) if (dir == D && reorder == RO && io == (IO == kernels::IO_TMA ? 0 : 1)) return run_cfg<13, T13::B, 1, 0, D, RO, IO, TW_LUT, T13::STAGES, 1, (IO == kernels::IO_TMA ? T13::PF : 0), A13>(i, o, n_ffts, grid, nullptr);
        C13(0, 1, kernels::IO_TMA) C13(0, 0, kernels::IO_TMA) C13(1, 1, kernels::IO_TMA) C13(1, 0, kernels::IO_TMA)
        C13(0, 1, kernels::IO_LDG) C13(0, 0, kernels::IO_LDG) C13(1, 1, kernels::IO_LDG) C13(1, 0, kernels::IO_LDG)
#undef C13
        return -3;
    }
    if (e == 14) {  // 16384 points (Tuning<14>): one 128 KB buffer, four TMA boxes per tile, [16,16,16,4]; which as for 8192, io 2 = TMA in / registers out
        using T14 = kernels::Tuning<14>;
        constexpr int A14 = kernels::ArithFor<14, 0, 1, 1>::value;
        const int reorder = which & 1, io = which >> 1;
#define C14(D, RO, IO) if (dir == D && reorder == RO && io == (IO == kernels::IO_TMA ? 0 : IO == kernels::IO_LDG ? 1 : 2)) return run_cfg<14, T14::B, 1, 0, D, RO, IO, TW_LUT, T14::STAGES, 1, (IO == kernels::IO_TMA ? T14::PF : 0), A14>(i, o, n_ffts, grid, nullptr);
        C14(0, 1, kernels::IO_TMA) C14(0, 0, kernels::IO_TMA) C14(1, 1, kernels::IO_TMA) C14(1, 0, kernels::IO_TMA)
        C14(0, 1, kernels::IO_LDG) C14(0, 0, kernels::IO_LDG) C14(1, 1, kernels::IO_LDG) C14(1, 0, kernels::IO_LDG)
        // io = 2: TMA in, registers out, the SAME buffer refilled behind the final exchange (hook_tail)
        C14(0, 1, kernels::IO_TMA_STG) C14(0, 0, kernels::IO_TMA_STG) C14(1, 1, kernels::IO_TMA_STG) C14(1, 0, kernels::IO_TMA_STG)
#undef C14
        return -3;
    }
    if (which == 2) {
        using T12 = kernels::TuningR32<12>;
        if (e == 12 && dir == 0) return run_cfg<12, T12::B, 1, 0, 0, 1, kernels::IO_TMA, TW_LUT, T12::STAGES, 1, T12::PF>(i, o, n_ffts, grid, nullptr);
        if (e == 12 && dir == 1) return run_cfg<12, T12::B, 1, 0, 1, 1, kernels::IO_TMA, TW_LUT, T12::STAGES, 1, T12::PF>(i, o, n_ffts, grid, nullptr);
        if (e == 5 && dir == 0) return run_cfg<5, 4, 128, 0, 0, 1, kernels::IO_TMA, TW_LUT, 3, 1, 1>(i, o, n_ffts, grid, nullptr);
        if (e == 5 && dir == 1) return run_cfg<5, 4, 128, 0, 1, 1, kernels::IO_TMA, TW_LUT, 3, 1, 1>(i, o, n_ffts, grid, nullptr);
        return -3;
    }

#define RA(E) if (e == E) return which == 0 ? run_reg_alt<E, 0>(i, o, n_ffts, dir, grid) : run_reg_alt<E, 1>(i, o, n_ffts, dir, grid);
    RA(7) RA(8) RA(9) RA(10)
#undef RA
    return -3;
}

int emu_alt_length(int variant)
{
    static const int n[] = {1024, 1024, 512, 128, 4096, 64, 32, 2048};
    return variant >= 0 && variant < 8 ? n[variant] : -1;
}

}
