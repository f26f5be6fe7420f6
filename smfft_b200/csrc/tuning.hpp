// tuning.hpp -- per-size kernel shapes of the native path (shared by the product build and the
// CPU SIMT-emulator tests).  E = log2 of the complex FFT length handled by one thread group.
//
//   B      log2 points per thread (R = 16: radix-16 register passes)
//   TILE_E log2 points per tile; a tile is F = 2^(TILE_E-E) whole FFTs, THREADS = 2^(TILE_E-B)
//   STAGES tile buffers per CTA on the TMA path (load k+1 / FFT k / store k-1 overlap)
//   MINB   CTAs per SM the kernel is compiled and launched for
// Values are the measured best of tools/tune (see profiles/); defaults before measurement came
// from the budget in SURVEY.md section 7.1.
#pragma once

namespace smfft {
namespace kernels {

template <int E>
struct Tuning {
    static constexpr int B = 4;
    static constexpr int TILE_E = E < 11 ? 11 : E;  // 2048-point tiles (16 KB), whole FFT for 4096
    static constexpr int F = 1 << (TILE_E - E);
    static constexpr int STAGES = 2;
    static constexpr int MINB = E == 12 ? 3 : 4;
};

}  // namespace kernels
}  // namespace smfft
