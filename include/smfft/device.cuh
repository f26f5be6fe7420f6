// smfft/device.cuh -- the NATIVE device primitive: a block FFT with 16 points per thread, callable inside user kernels.
//
// KAdamek/SMFFT exists to be called from inside other kernels (README.md:2, 10-14 of the reference: "the FFT ... can be
// used as a part of a kernel"; its device entry point is do_SMFFT_CT_DIT, SMFFT_CooleyTukey_C2C/FFT-GPU-32bit.cu:334-532).
// include/smfft/compat.cuh keeps that entry point and its contract (4 points per thread, tile in shared memory).  This
// header is what a NEW caller on B200 should use instead: the same block FFT the library's own kernels run
// (detail/block_fft.cuh: Stockham autosort, register-resident radix-16 passes, bank-conflict-free swizzled exchanges),
// with the data kept in REGISTERS across the call, so that work placed between two transforms -- the pointwise multiply
// of a convolution -- costs no trip through shared memory and no extra barrier.
//
//   using F = smfft::BlockFFT<10, smfft::FORWARD>;            // 1024 points, 64 threads per transform
//   using I = smfft::BlockFFT<10, smfft::INVERSE>;
//   __global__ void __launch_bounds__(F::THREADS) convolve(const float2* x, const float2* H, float2* y) {
//       __shared__ __align__(16) float2 xch[F::EXCHANGE_POINTS];   // scratch for the exchanges between passes
//       float2 v[F::R];
//       F::load(v, x + (size_t)blockIdx.x * F::TILE_POINTS);       // coalesced: v[m] = x[F::index(m)]
//       F::exec(v, xch);                                           // v[m] = X[F::index(m)]
//       for (int m = 0; m < F::R; m++) v[m] = mul(v[m], H[F::index(m) % F::N]);   // fused work, in registers
//       I::exec(v, xch);                                           // same ownership in and out: no re-layout
//       I::store(v, y + (size_t)blockIdx.x * F::TILE_POINTS);
//   }
//   convolve<<<n_transforms / F::FFTS, F::THREADS>>>(x, H, y);
// (smfft::block_convolve below is exactly this with the pointwise step as a functor; tests/compat/compat_kernels.cu and
// tools/convolve_bench.py use and time it against the reference's own device function in the same user kernel.)
// The library's own two-pass transforms of 2^15 .. 2^18 points (smfft_b200/csrc/big_fft.cu) are user kernels of this primitive
// too: 16 transforms per block read from a TMA-loaded tile, exec(), a twiddle multiply on the registers, TMA store.
//
// CONTRACT
//   * blockDim.x == THREADS = FFTS * N / R (1-D block), every thread of the block calls exec() (it contains __syncthreads);
//   * ownership: thread t of transform f (f = threadIdx.x / T, t = threadIdx.x % T, T = N / R) holds
//     v[m] = x[f*N + t + m*T], m = 0..R-1, on entry AND on exit (natural order, un-normalised, sign -/+ for
//     FORWARD/INVERSE like the reference); index(m) returns f*N + t + m*T;
//   * `exchange` is shared memory, 16-byte aligned, EXCHANGE_POINTS float2, contents undefined afterwards; the same buffer
//     may be passed to the next exec() without a barrier in between (each exec synchronises before its first write);
//   * twiddles: TW_MUFU (default; __sincosf like the reference, nothing to set up) or TW_LUT (pass `tw`, a shared-memory
//     table of TWIDDLE_POINTS float2 filled once per block by fill_twiddles() from the global table whose device address
//     smfft_twiddle_table() of the C ABI returns; one accurate base twiddle per pass, powers in registers);
//   * N = 2^LOG2N, 32 <= N <= 4096 (one transform never spans blocks, as in the reference); R = 16, or 32 with LOG2R = 5;
//     FFTS (transforms per block) a power of two -- use it to give small transforms enough threads per block.
#pragma once
#include "detail/block_fft.cuh"

namespace smfft {

enum Direction { FORWARD = 0, INVERSE = 1 };

template <int LOG2N, int DIR, int FFTS_PER_BLOCK = 1, int TW = TW_MUFU, int LOG2R = 4, int PACKED = 0>
struct BlockFFT {
    static_assert(LOG2N >= 5 && LOG2N <= 12, "32 .. 4096 points");
    static_assert(LOG2R == 4 || LOG2R == 5, "16 or 32 points per thread");
    // exchanges in the swizzled layout of the library's kernels; R = 32 spreads a thread's 32 contiguous first-pass outputs with SW256
    using XL = typename std::conditional<LOG2R == 5, detail::LayoutSW256, detail::LayoutSW128>::type;
    using Cfg = detail::BlockCfg<LOG2N, LOG2R, FFTS_PER_BLOCK, DIR, 1, TW, detail::LayoutSW128, XL, true, true, PACKED ? 2 : 0>;

    static constexpr int N = Cfg::N;              // points per transform
    static constexpr int R = Cfg::R;              // points per thread
    static constexpr int T = Cfg::T;              // threads per transform
    static constexpr int FFTS = FFTS_PER_BLOCK;   // transforms per block
    static constexpr int THREADS = Cfg::THREADS;  // = blockDim.x
    static constexpr int TILE_POINTS = Cfg::L;    // points per block
    static constexpr int EXCHANGE_POINTS = Cfg::L;
    static constexpr int TWIDDLE_POINTS = TW == TW_LUT ? (Cfg::TW_C2C_ENTRIES > 0 ? Cfg::TW_C2C_ENTRIES : 1) : 0;

    // element (within the block's tile) held by register m of the calling thread
    static __device__ __forceinline__ int index(int m) { return ((threadIdx.x >> Cfg::A) << Cfg::E) + (threadIdx.x & (T - 1)) + m * T; }

    // v[m] = tile[index(m)]: consecutive threads read consecutive points (global memory: coalesced; shared memory: conflict-free)
    static __device__ __forceinline__ void load(float2 (&v)[R], const float2* __restrict__ tile)
    {
        const int x0 = index(0);
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = tile[x0 + m * T];
    }
    static __device__ __forceinline__ void store(const float2 (&v)[R], float2* __restrict__ tile)
    {
        const int x0 = index(0);
#pragma unroll
        for (int m = 0; m < R; m++) tile[x0 + m * T] = v[m];
    }

    // TW_LUT only: fill the block's twiddle table from the library's global table (smfft_twiddle_table()); synchronise before the first exec
    static __device__ __forceinline__ void fill_twiddles(float2* tw, const float2* __restrict__ wtable)
    {
        detail::fill_twiddle_table<Cfg, false, 0>(tw, wtable, threadIdx.x, THREADS);
    }

    // the transform, registers to registers
    static __device__ __forceinline__ void exec(float2 (&v)[R], float2* exchange, const float2* tw = nullptr)
    {
        const int t = threadIdx.x & (T - 1);
        const int fbase = (threadIdx.x >> Cfg::A) << Cfg::E;
        detail::run_passes<Cfg, 0, detail::XF_C2C>(v, exchange, fbase, t, t, tw, detail::NoHook{});
    }
};

// forward transform, pointwise functor, inverse transform -- all on the same registers.
// pointwise(value, k) -> value is applied to every spectrum element X[k], k = index within the transform (0..N-1).
template <class FWD, class INV, class Pointwise>
__device__ __forceinline__ void block_convolve(float2 (&v)[FWD::R], float2* exchange, Pointwise&& pointwise, const float2* tw_fwd = nullptr,
                                               const float2* tw_inv = nullptr)
{
    static_assert(FWD::N == INV::N && FWD::R == INV::R && FWD::FFTS == INV::FFTS, "forward and inverse must share the shape");
    FWD::exec(v, exchange, tw_fwd);
#pragma unroll
    for (int m = 0; m < FWD::R; m++) v[m] = pointwise(v[m], FWD::index(m) & (FWD::N - 1));
    INV::exec(v, exchange, tw_inv);
}

}  // namespace smfft
