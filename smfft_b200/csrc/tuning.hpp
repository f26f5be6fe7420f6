// tuning.hpp -- per-size kernel shapes of the native path (shared by the product build and the
// CPU SIMT-emulator tests).  E = log2 of the complex FFT length handled by one thread group.
//
//   B      log2 points per thread (R = 16: radix-16 register passes)
//   TILE_E log2 points per tile; a tile is F = 2^(TILE_E-E) whole FFTs, THREADS = 2^(TILE_E-B)
//   STAGES tile buffers per CTA on the TMA paths (load k+1 overlaps the FFT of tile k)
//   MINB   CTAs per SM the kernel is compiled for (__launch_bounds__ register cap)
//   CTAS   CTAs per SM the persistent grid is launched with
//   STG / STG_R2C / STG_C2R   1 = results leave from registers (IO_TMA_STG), 0 = TMA stores, per transform kind
//          (register stores need T >= 16 threads per FFT to stay coalesced: not for N <= 128)
//
// What the measurements say (profiles/r01_tune_*.csv, profiles/r01_copylab_*.csv, B200, 4 GiB batch):
//   * the FFT arithmetic is fully hidden: a staging-only kernel (tile in, tile out) costs the same;
//   * the memory system wants 48-64 KB of outstanding TMA loads per SM: CTAS x (STAGES-1) x tile
//     bytes.  32 KB is too little (1.4-1.5 ms), >= 96 KB is too much (1.36-1.40 ms), the sweet spot
//     gives 1.27-1.30 ms for a pure copy and 1.28-1.32 ms for the FFT (cudaMemcpy D2D: 1.30 ms);
//   * hence two CTAs of 256 threads with two 32 KB buffers each (three CTAs x 16 KB for N <= 64).
#pragma once

namespace smfft {
namespace kernels {

template <int E>
struct Tuning {  // N >= 1024
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - E), STAGES = 2, MINB = 2, CTAS = 2, STG = 0, STG_R2C = 1, STG_C2R = 0;
};
template <>
struct Tuning<5> {
    static constexpr int B = 4, TILE_E = 11, F = 1 << (TILE_E - 5), STAGES = 2, MINB = 4, CTAS = 3, STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<6> {
    static constexpr int B = 4, TILE_E = 11, F = 1 << (TILE_E - 6), STAGES = 2, MINB = 4, CTAS = 3, STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<7> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 7), STAGES = 2, MINB = 2, CTAS = 2, STG = 1, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<8> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 8), STAGES = 2, MINB = 2, CTAS = 2, STG = 1, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<9> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 9), STAGES = 2, MINB = 2, CTAS = 2, STG = 1, STG_R2C = 1, STG_C2R = 0;
};

}  // namespace kernels
}  // namespace smfft
