// tuning.hpp -- per-size kernel shapes of the native path (shared by the product build and the
// CPU SIMT-emulator tests).  E = log2 of the complex FFT length handled by one thread group.
//
//   B      log2 points per thread (R = 16: radix-16 register passes)
//   TILE_E log2 points per tile; a tile is F = 2^(TILE_E-E) whole FFTs, THREADS = 2^(TILE_E-B)
//   STAGES tile buffers per CTA on the TMA path (load k+1 / FFT k / store k-1 overlap)
//   MINB   CTAs per SM the kernel is compiled for (__launch_bounds__)
// Values are the measured best of tools/tune on a B200, 4 GiB batch (profiles/r01_tune_shapes_a.csv:
// 450 variants, every one checked against the default shape's output).  All shapes within ~3 % of
// each other at 1.32-1.42 ms; the LDG-staged variants of the same shapes take 2.1-3.1 ms.
#pragma once

namespace smfft {
namespace kernels {

template <int E>
struct Tuning {  // N = 256 and below: 4 CTAs/SM x one 32 KB tile buffer
    static constexpr int B = 4;
    static constexpr int TILE_E = 12;
    static constexpr int F = 1 << (TILE_E - E);
    static constexpr int STAGES = 1;
    static constexpr int MINB = 4;
};
template <>
struct Tuning<5> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 5), STAGES = 2, MINB = 2;
};
template <>
struct Tuning<7> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 7), STAGES = 2, MINB = 2;
};
template <>
struct Tuning<9> {
    static constexpr int B = 4, TILE_E = 11, F = 1 << (TILE_E - 9), STAGES = 2, MINB = 4;
};
template <>
struct Tuning<10> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 10), STAGES = 3, MINB = 2;
};
template <>
struct Tuning<11> {
    static constexpr int B = 4, TILE_E = 12, F = 1 << (TILE_E - 11), STAGES = 3, MINB = 2;
};
template <>
struct Tuning<12> {
    static constexpr int B = 4, TILE_E = 12, F = 1, STAGES = 2, MINB = 3;
};

}  // namespace kernels
}  // namespace smfft
