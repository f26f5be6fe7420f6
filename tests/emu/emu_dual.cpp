// emu_dual.cpp -- the dual-lane shapes on the SIMT emulator (own translation unit: compiles in parallel).  TESTS ONLY.
#include "emu_run_cfg.hpp"

// dual-lane transforms (block_fft_dual.cuh): two transforms per thread in the packed f32x2 lanes.
// kind: 0 C2C fwd natural, 1 C2C inv no-reorder, 2 C2C fwd no-reorder, 3 C2C inv natural, 4 R2C, 5 C2R; io as kernels::IO_*
template <int E, int B, int F, int STAGES, int PF>
static int run_dual_shape(const float2* i, float2* o, long long n, int kind, int io, int tw, int reps, int grid, double* bf)
{
    using namespace kernels;
#define DC(K, MODE, D, RO, I, T, RP)                                                                    \
    if (kind == K && io == I && tw == T && reps == RP)                                                  \
        return run_cfg<E, B, F, MODE, D, RO, I, T, STAGES, RP, (I == IO_TMA ? PF : (PF < 0 ? 0 : PF)), 1>(i, o, n, grid, bf);
    DC(0, 0, 0, 1, IO_TMA, 0, 1) DC(1, 0, 1, 0, IO_TMA, 0, 1) DC(2, 0, 0, 0, IO_TMA, 0, 1) DC(3, 0, 1, 1, IO_TMA, 0, 1)
    DC(4, 1, 0, 1, IO_TMA, 0, 1) DC(5, 2, 1, 1, IO_TMA, 0, 1)
    DC(0, 0, 0, 1, IO_TMA_STG, 0, 1) DC(1, 0, 1, 0, IO_TMA_STG, 0, 1) DC(4, 1, 0, 1, IO_TMA_STG, 0, 1) DC(5, 2, 1, 1, IO_TMA_STG, 0, 1)
    DC(0, 0, 0, 1, IO_LDG, 0, 1) DC(2, 0, 0, 0, IO_LDG, 0, 1) DC(4, 1, 0, 1, IO_LDG, 0, 1) DC(5, 2, 1, 1, IO_LDG, 0, 1)
    DC(0, 0, 0, 1, IO_TMA, 1, 1) DC(4, 1, 0, 1, IO_TMA, 1, 1) DC(5, 2, 1, 1, IO_LDG, 1, 1)
    DC(0, 0, 0, 1, IO_LDG, 0, 3) DC(2, 0, 0, 0, IO_LDG, 0, 3) DC(4, 1, 0, 1, IO_LDG, 0, 3)
#undef DC
    return -1;
}

extern "C" {

int emu_run_dual(const void* in, void* out, int e, int b, long long n_ffts, int kind, int io, int tw, int reps, int grid,
                 double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    if (b == 4) {
        switch (e) {
            case 8: return run_dual_shape<8, 4, 16, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 9: return run_dual_shape<9, 4, 8, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 10: return run_dual_shape<10, 4, 4, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 11: return run_dual_shape<11, 4, 2, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 12: return run_dual_shape<12, 4, 2, 2, 2>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
        }
    } else if (b == 5) {
        switch (e) {
            case 9: return run_dual_shape<9, 5, 8, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 10: return run_dual_shape<10, 5, 4, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 11: return run_dual_shape<11, 5, 2, 3, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
        }
    }
    return -1;
}

}
