// tuning.hpp -- per-size kernel shapes of the native path (shared by the product build and the
// CPU SIMT-emulator tests).  E = log2 of the complex FFT length handled by one thread group.
//
//   B      log2 points per thread (R = 16: radix-16 register passes)
//   TILE_E log2 points per tile; a tile is F = 2^(TILE_E-E) whole FFTs, THREADS = 2^(TILE_E-B)
//   STAGES tile buffers per CTA on the TMA paths (the load of tile k+1 overlaps the FFT of tile k)
//   MINB   CTAs per SM the kernel is compiled for (__launch_bounds__ register cap)
//   CTAS   CTAs per SM the persistent grid is launched with
//   PF     pass after which the next tile's load is issued (-1: top of the iteration; TMA-store kernels only)
//   STG / STG_R2C / STG_C2R   1 = results leave from registers (IO_TMA_STG), 0 = TMA stores, per transform kind
//          (register stores need T >= 16 threads per FFT to stay coalesced: not for N <= 128)
//
// What the measurements say (profiles/r01_tune_*.csv, profiles/r01_copylab_*.csv, B200, 4 GiB batch):
//   * the FFT arithmetic is hidden at N <= 512: a staging-only kernel (tile in, tile out) costs the same;
//   * the memory system wants 48-64 KB of outstanding TMA loads per SM.  32 KB is too little
//     (1.4-1.5 ms), >= 96 KB is too much (1.36-1.40 ms), the sweet spot gives 1.27-1.30 ms for a pure
//     copy and 1.28-1.32 ms for the FFT (cudaMemcpy D2D: 1.30 ms);
//   * N >= 1024 (and R2C/C2R) are close to issue-bound (49-60 SASS instructions per point, 8 warps per
//     32 KB tile): they want more resident warps than "2 CTAs x one prefetched tile" allows, so the
//     prefetch is issued late (after pass 1) and 3-8 smaller CTAs share the same load concurrency.
#pragma once
#include <type_traits>

namespace smfft {
namespace kernels {

template <int E>
struct Tuning;
// N = 32: [TILE 12, 3 stages, 1 CTA/SM] is ~1 % faster at full clocks (1.26-1.30 ms) but has only 8 warps per SM and
// scales with 1/clock: 1.43-1.46 ms at the 1.7 GHz a power-capped run settles at.  Six small CTAs hold 1.29 ms there.
template <>
struct Tuning<5> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 5), STAGES = 2, MINB = 8, CTAS = 6, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<6> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 6), STAGES = 2, MINB = 8, CTAS = 6, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<7> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 7), STAGES = 2, MINB = 8, CTAS = 6, PF = -1;
    static constexpr int STG = 1, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<8> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 8), STAGES = 2, MINB = 8, CTAS = 6, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct Tuning<9> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 9), STAGES = 2, MINB = 8, CTAS = 8, PF = 1;
    static constexpr int STG = 0, STG_R2C = 1, STG_C2R = 0;
};
template <>
struct Tuning<10> {
    static constexpr int B = 4, TILE_E = 10, F = 1, STAGES = 2, MINB = 8, CTAS = 8, PF = 1;
    static constexpr int STG = 0, STG_R2C = 1, STG_C2R = 0;
};
template <>
struct Tuning<11> {
    static constexpr int B = 4, TILE_E = 11, F = 1, STAGES = 2, MINB = 6, CTAS = 5, PF = 1;
    static constexpr int STG = 0, STG_R2C = 1, STG_C2R = 0;
};
#ifndef SMFFT_T12_STAGES
#define SMFFT_T12_STAGES 2
#define SMFFT_T12_MINB 3
#define SMFFT_T12_CTAS 3
#define SMFFT_T12_STG 0
#endif
template <>
struct Tuning<12> {
    static constexpr int B = 4, TILE_E = 12, F = 1, STAGES = SMFFT_T12_STAGES, MINB = SMFFT_T12_MINB, CTAS = SMFFT_T12_CTAS, PF = 1;
    static constexpr int STG = SMFFT_T12_STG, STG_R2C = 1, STG_C2R = 0;
};

// 8192 points (beyond the reference's range): one transform still lives in one CTA's shared memory -- a
// 64 KB tile, three stages, R = 32 ([32,32,8]: 256 threads), one CTA per SM -- so the reference's premise (one FFT never
// leaves shared memory) holds one size further.  C2C only, both orders, both directions.
#ifndef SMFFT_T13_STAGES
#define SMFFT_T13_STAGES 3  // three 64 KB buffers: 1.375 / 1.57 ms against 1.59 / 1.78 ms with two (profiles/r02_ab_8192_shapes.json); R = 16 (512 threads, four passes): 1.68 / 1.77 ms
#define SMFFT_T13_MINB 1
#define SMFFT_T13_CTAS 1
#define SMFFT_T13_STG 0
#define SMFFT_T13_B 5
#endif
template <>
struct Tuning<13> {
    static constexpr int B = SMFFT_T13_B, TILE_E = 13, F = 1, STAGES = SMFFT_T13_STAGES, MINB = SMFFT_T13_MINB, CTAS = SMFFT_T13_CTAS, PF = 1;
    static constexpr int STG = SMFFT_T13_STG, STG_R2C = 0, STG_C2R = 0;
};

// 16384 points: the tile is 128 KB, so ONE buffer per CTA and one CTA per SM.  R = 32, [32,32,16]: 512 threads (16 warps,
// 112 registers), two exchanges.  The results leave from registers (IO_TMA_STG) and the refill of the one buffer is
// issued behind the final exchange, so the next tile's load overlaps the last pass and the stores: 2.43 / 2.36 ms (R = 16,
// [16,16,16,4], TMA store, load -> transform -> store in sequence) -> 1.99 / 1.91 ms (same plan, registers out) -> 1.76 / 1.78 ms
// (R = 32), natural / bit-reversed input, 4 GiB batches (profiles/r02_ab_16384_single_buffer.json).  What is left is not a
// wait for memory (ncu: no long-scoreboard stalls) but 16 warps alone on an SM between block barriers (issue slots 43 %
// busy): 0.74 of the roofline is the price of keeping one transform inside one CTA's shared memory.  C2C only.
#ifndef SMFFT_T14_STG
#define SMFFT_T14_STG 1
#endif
#ifndef SMFFT_T14_B
#define SMFFT_T14_B 5  // reversed [4,16,16,16]: 2.14 / 2.05 ms; scalar arithmetic or MUFU twiddles: slower in one order or the other
#endif
#ifndef SMFFT_T14_ARITH
#define SMFFT_T14_ARITH 2
#endif
template <>
struct Tuning<14> {
    static constexpr int B = SMFFT_T14_B, TILE_E = 14, F = 1, STAGES = 1, MINB = 1, CTAS = 1, PF = -1;
    static constexpr int STG = SMFFT_T14_STG, STG_R2C = 0, STG_C2R = 0;
};

// Natural-order transforms of 512, 1024 and 4096 points (CT reorder=1, Stockham) run R = 32
// points per thread: [32,16] / [32,32] needs ONE exchange instead of two -- 39-44 SASS instructions per
// point instead of 48-52, 48 instead of 64 bytes of shared-memory traffic per point
// (profiles/r01_tune_radix32_r.csv: 1.286 ms vs 1.31-1.33 ms).  The no-reorder transform stays on R = 16:
// its first read is a contiguous run per thread, which for 32 points spans two rows and conflicts 2-way.
template <int E>
struct TuningR32 {
    static constexpr int B = 5, TILE_E = 12, F = 1 << (TILE_E - E), STAGES = 2, MINB = 2, CTAS = 2, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
// 4096 = [32,32,4]: still two exchanges, but 128 threads carry the tile instead of 256 (profiles/r01_tune_c2c_r32.csv)
template <>
struct TuningR32<12> {
    static constexpr int B = 5, TILE_E = 12, F = 1, STAGES = 2, MINB = 3, CTAS = 3, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};

// R2C / C2R (external, REPS == 1) have their own shapes: the real pass makes them the most issue-heavy
// kernels of the library, so they want more resident warps per byte in flight than the C2C kernels, and the
// 512 / 1024-point cores run R = 32 (profiles/r01_tune_real_shapes.csv, sustained sweep of both kinds).
template <int E>
struct TuningReal;
template <>
struct TuningReal<5> {
    static constexpr int B = 4, TILE_E = 11, F = 1 << (TILE_E - 5), STAGES = 2, MINB = 6, CTAS = 4, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct TuningReal<6> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 6), STAGES = 2, MINB = 8, CTAS = 8, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct TuningReal<7> {
    static constexpr int B = 4, TILE_E = 10, F = 1 << (TILE_E - 7), STAGES = 2, MINB = 8, CTAS = 8, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct TuningReal<8> {
    static constexpr int B = 4, TILE_E = 11, F = 1 << (TILE_E - 8), STAGES = 2, MINB = 6, CTAS = 4, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
template <>
struct TuningReal<9> {
    static constexpr int B = 5, TILE_E = 11, F = 1 << (TILE_E - 9), STAGES = 2, MINB = 4, CTAS = 4, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 1;
};
template <>
struct TuningReal<10> {
    static constexpr int B = 5, TILE_E = 11, F = 1 << (TILE_E - 10), STAGES = 2, MINB = 4, CTAS = 4, PF = 1;
    static constexpr int STG = 0, STG_R2C = 1, STG_C2R = 1;
};
// experiment switches (variant builds): shape of the 4096-real kernels.  One buffer and 8 or 10 CTAs per SM (32 / 40 warps, the CTAs
// interleave instead of each pipelining two tiles) against the product's two buffers x 6 CTAs: R2C -1..-3 %, C2R +3..+4 %
// (profiles/r02_ab_real4096_single_stage.json) -- inside the box-to-box spread, the product shape stays.
#ifndef SMFFT_TR11_STAGES
#define SMFFT_TR11_STAGES 2
#define SMFFT_TR11_MINB 6
#define SMFFT_TR11_CTAS 6
#define SMFFT_TR11_PF 1
#define SMFFT_TR11_STG_R2C 1
#endif
template <>
struct TuningReal<11> {
    static constexpr int B = 4, TILE_E = 11, F = 1, STAGES = SMFFT_TR11_STAGES, MINB = SMFFT_TR11_MINB, CTAS = SMFFT_TR11_CTAS, PF = SMFFT_TR11_PF;
    static constexpr int STG = 0, STG_R2C = SMFFT_TR11_STG_R2C, STG_C2R = 0;  // mirrored R2C: the descending half of the result leaves from registers
};
template <>
struct TuningReal<12> {
    static constexpr int B = 4, TILE_E = 12, F = 1, STAGES = 2, MINB = 3, CTAS = 3, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
// 16384 reals (beyond the reference): the 8192-point core, the shape of Tuning<13> -- R = 32 [32,32,8] (C2R: reversed [8,32,32]),
// three 64 KB buffers, one CTA per SM; the mirrored real passes own their pairs (U = 4 butterflies per thread in the last pass)
template <>
struct TuningReal<13> {
    static constexpr int B = 5, TILE_E = 13, F = 1, STAGES = 3, MINB = 1, CTAS = 1, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};

// arithmetic flavour of one kernel instance (BlockCfg::DUAL_, bit flags): 0 = scalar, 1 = dual-lane, 2 = packed (re, im)
// add / subtract in the butterflies, 4 = reversed pass plan.  Packed add / subtract removes ~9 % of the issue slots; it pays where a kernel is issue-bound
// (interleaved A/B against the all-scalar build, profiles/r01_ab_scalar_vs_product.json and ..._vs_packed_all.json:
// C2R -1.4..-7.7 % at every size, R2C of 4096 / 8192 reals -4..-6 %, FFT_multiple -2..-11 %) and is neutral or slightly
// negative (+-1 %) for the HBM-bound C2C external kernels and the other R2C sizes, which stay scalar; the R = 32 R2C
// FFT_multiple shapes (512 / 1024 points) lose 1-3 % and stay scalar too.  Under the sustained load of the 16-launch
// bench step (power-capped clocks) the three-pass R = 16 C2C kernels gain as well -- 2048 points -3..-5 %, 4096 points
// no-reorder -4 % (profiles/r01_bench_sustained_scalar_vs_packed.json) -- the R = 32 shapes do not (+1 %).
// The dual-lane form (block_fft_dual.cuh: two transforms per thread, ALL arithmetic and exchanges packed, half the
// instructions per point) is NOT used by the product: it also halves the resident warps, and the large and real
// kernels turn out to be bound by shared-memory wavefronts and latency, not by issue slots alone
// (profiles/r01_tune_dual_a.csv: equal at 2048 points, 3-12 % slower elsewhere).  It stays as a measured experiment
// with emulator coverage (tools/tune_dual, tests/test_emu_kernels.py).
// FFT_multiple of 32 points (2 lanes per transform, plan [16, 2]): the one exchange goes through warp shuffles (flag 8) --
// these kernels run the shared-memory pipe at 81-83 % (profiles/roofline_traffic.json, "multiple"), and the shuffle form
// moves each value once instead of writing and reading it.  Interleaved A/B (profiles/r02_ab_xshfl_multiple.json):
// 32 points 0.571 -> 0.444 ms (-22 %); 64 points (4 lanes: a 4x4 transposition per four registers, 24 SHFL + 128 SEL against
// 8 STS.128 + 16 LDS.64) +7 % / -1 %: the selects cost what the wavefronts save, so it stays on shared memory.
#ifndef SMFFT_XSHFL_MULTIPLE
#define SMFFT_XSHFL_MULTIPLE 1
#endif
#if defined(SMFFT_R32_E12)
constexpr bool kNoR32E12 = false;
#else
constexpr bool kNoR32E12 = true;
#endif
template <int E, int MODE, int REORDER, int REPS>
struct ArithFor {
#if defined(SMFFT_FORCE_ARITH)
    static constexpr int value = SMFFT_FORCE_ARITH;
#else
    static constexpr int value = REPS > 1 ? ((MODE == 1 && (E == 9 || E == 10)) ? 0 : (MODE == 0 && E == 5 && SMFFT_XSHFL_MULTIPLE) ? 10 : 2)
                                 : (MODE == 2 && (E == 11 || E == 12 || E == 13)) ? 6  // + reversed plan: the C2R pass owns its pairs (MirrorC2R)
                                 : (MODE == 2 || (MODE == 1 && (E >= 11 || E == 5))) ? 2
                                 : (MODE == 0 && E == 14) ? SMFFT_T14_ARITH
                                 : (MODE == 0 && (E == 11 || E == 13 || (E == 12 && (REORDER == 0 || kNoR32E12)))) ? 2 : 0;  // C2C on R = 16 plans (and 8192 points), sustained load
#endif
};

// R2C of 8192 reals: the R = 32 plan [32,32,4] runs U = 8 butterflies per thread in its last pass, so the real pass owns
// its pairs (MirrorR2C with four mirror pairs per thread): 1.59 -> 1.50 ms against the R = 16 shape, whose last pass has
// U = 1.  (2048 reals: the same trick on [16,16,4] only ties with the single-exchange R = 32 shape, 1.332 vs 1.331 ms.)
struct TuningR2C12 {
    static constexpr int B = 5, TILE_E = 12, F = 1, STAGES = 2, MINB = 3, CTAS = 3, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
// C2R of 8192 reals: the mirror image, reversed R = 32 plan [4,32,32] (MirrorC2R with four mirror pairs per thread):
// 1.59 -> 1.49 ms.  (2048 reals on the reversed R = 16 plan [4,16,16] loses to the single-exchange shape, 1.39-1.44 vs 1.35 ms.)
struct TuningC2R12 {
    static constexpr int B = 5, TILE_E = 12, F = 1, STAGES = 2, MINB = 3, CTAS = 3, PF = 1;
    static constexpr int STG = 0, STG_R2C = 0, STG_C2R = 0;
};
// Register-direct input (IO_REG) for natural-order C2C: global -> registers -> passes -> global, shared memory only for the
// exchange.  With the driver's DEFAULT L1 carve-out (not max shared) the LDG/STG path is as fast as cuFFT's kernels
// (tools/fftlike_copy.cu): 1.252 ms at 1024 points against 1.29-1.30 ms for the TMA path
// (profiles/r01_tune_register_direct_carveout_default.csv).  PREFER follows the sustained bench step, and there the
// shape loses even with 16 warps per SM: 1.257 vs 1.303 ms in single launches but 1.40 vs 1.34 ms inside the power-capped step
// (profiles/r01_register_direct_1024_burst_vs_sustained.json).  The instance stays reachable with io = 4.
// Round 2: cuFFT's own kernels for 128..1024 points have exactly this structure (profiles/r02_cufft_kernel_shapes.csv:
// vector_fft<N, EPT<16|32>>, 64..128 threads, 4..12 transforms per CTA, 63..121 registers, one CTA per tile, 4..17 KB of
// shared memory for the one exchange, large L1) and run 1.23-1.26 ms against 1.29-1.30 ms for the TMA kernels in the same
// sustained step (profiles/r02_bench_c.json).  Every size from 128 to 1024 points therefore carries up to two
// register-direct shapes as ALTERNATES: reachable with io = 4 (shape A) / io = 5 (shape B), and candidates of the
// first-use selection (option "select" = 1, launch.cu), which times them against the default instance on the caller's own
// batch and keeps the fastest -- the measured regime (burst or sustained, this box's power cap) picks the shape.
template <int E>
struct RegDirect {
    static constexpr bool ON = false, PREFER = false, ON_B = false;
    static constexpr int B = 4, TILE_E = 10, MINB = 8;
    static constexpr int B_B = 4, TILE_E_B = 10, MINB_B = 8;
};
template <>
struct RegDirect<7> {  // cuFFT: 16 points per thread, 8 threads per transform, 63 registers
    static constexpr bool ON = true, PREFER = false, ON_B = true;
    static constexpr int B = 4, TILE_E = 10, MINB = 16;       // A: 8 transforms per 64-thread CTA, <= 64 registers
    static constexpr int B_B = 4, TILE_E_B = 11, MINB_B = 6;  // B: 16 transforms per 128-thread CTA
};
// 256 points is the one size where the register-direct shape wins in EVERY regime measured (profiles/r02_select_probe_j.json:
// 1.248-1.258 ms against 1.297 ms for the TMA instance inside the interleaved sustained step, 1.270 vs 1.306 ms in its own
// consecutive steps, 1.241 vs 1.298 ms in short bursts; cuFFT 1.33-1.36 ms): it is the static default there.
template <>
struct RegDirect<8> {  // cuFFT: 16 points per thread, 4 transforms per 64-thread CTA, 80 registers
    static constexpr bool ON = true, PREFER = true, ON_B = true;
    static constexpr int B = 4, TILE_E = 10, MINB = 12;
    static constexpr int B_B = 4, TILE_E_B = 11, MINB_B = 6;
};
template <>
struct RegDirect<9> {  // cuFFT: 32 points per thread, 4 transforms per 64-thread CTA, 121 registers
    static constexpr bool ON = true, PREFER = false, ON_B = true;
    static constexpr int B = 5, TILE_E = 11, MINB = 8;
    static constexpr int B_B = 5, TILE_E_B = 10, MINB_B = 16;  // B: ONE WARP per CTA (two transforms): block barriers cost nothing, 16 CTAs per SM
};
template <>
struct RegDirect<10> {
    static constexpr bool ON = true, ON_B = true;
#if defined(SMFFT_REGDIRECT_DEFAULT)
    static constexpr bool PREFER = true;
#else
    static constexpr bool PREFER = false;
#endif
    static constexpr int B = 5, TILE_E = 11, MINB = 8;        // A: 93 registers: eight CTAs = 16 warps per SM (six: 1.263 / 1.42 ms)
    static constexpr int B_B = 5, TILE_E_B = 10, MINB_B = 16;  // B: ONE WARP per CTA = one transform (cuFFT's 4-transform / 128-thread shape measured 1.29-1.33 ms)
};

// shape of one kernel instance: MODE 0 C2C / 1 R2C / 2 C2R (kernels::MODE_*), REPS > 1 = FFT_multiple
// (compute-bound: R = 32 pays at 512 and 1024 points, C2C 0.88 / 1.09 ms vs 1.12 / 1.13 ms, R2C 1.21 / 1.23 vs
// 1.28 / 1.27 ms, not at 4096)
template <int E, int MODE, int REORDER, int REPS>
struct ShapeFor {
    // 4096 points: the R = 32 plan [32,32,4] wins single launches (1.335 vs 1.36 ms) but runs 12 warps per SM and loses
    // under the sustained, power-capped load of the bench step (1.47 vs 1.42 ms for R = 16 with packed add / subtract,
    // profiles/r01_bench_sustained_r32_vs_r16_4096.json); SMFFT_R32_E12 brings it back for experiments.  The same test for
    // 512 / 1024 points keeps R = 32 (profiles/r01_bench_sustained_r16_512_1024.json: 6462 vs 6435 GB/s for the step)
#if defined(SMFFT_R32_E12)
    static constexpr bool R32 = REORDER == 1 && ((MODE == 0 && (E == 9 || E == 10 || (E == 12 && REPS == 1))) ||
#else
    static constexpr bool R32 = REORDER == 1 && ((MODE == 0 && (E == 9 || E == 10)) ||
#endif
                                                 (MODE != 0 && REPS > 1 && (E == 9 || E == 10)));
    static constexpr bool REAL = MODE != 0 && REPS == 1;
#if defined(SMFFT_R2C11_R32)
    // experiment: R2C of 4096 reals on the mirrored R = 32 plan [32,32,2] (12 warps per SM).  Measured (profiles/r02_ab_r2c4096_r32.json):
    // 1.37 ms in the first launches after idle, but 1.63 vs 1.49 ms interleaved and 1.60 vs 1.53 ms in its own steady state -- not used.
    static constexpr bool M12 = MODE == 1 && REPS == 1 && (E == 12 || E == 11);
#else
    static constexpr bool M12 = MODE == 1 && REPS == 1 && E == 12;
#endif
    static constexpr bool C12 = MODE == 2 && REPS == 1 && E == 12;
    using type = typename std::conditional<M12, TuningR2C12, typename std::conditional<C12, TuningC2R12,
                 typename std::conditional<REAL, TuningReal<E>, typename std::conditional<R32, TuningR32<E>, Tuning<E>>::type>::type>::type>::type;
};

}  // namespace kernels
}  // namespace smfft
