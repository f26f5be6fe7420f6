// smfft/compat.cuh -- the reference's DEVICE API, re-implemented on the native block FFT.
//
// Drop-in for what KAdamek/SMFFT exposes to user kernels (README.md:10-20, 56-60):
//   trait classes      FFT_Params, FFT_{32..4096}_{forward,inverse}[_noreorder]
//                                              SMFFT_CooleyTukey_C2C/SM_FFT_parameters.cuh:1-390
//                      FFT_ConstParams, FFT_{256..4096}, FFT_ConstDirection, FFT_forward, FFT_inverse
//                                              SMFFT_Stockham_R2C_C2R/FFT-GPU-32bit-Stockham.cu:15-81
//   device functions   do_SMFFT_CT_DIT<P>                 CT/FFT-GPU-32bit.cu:334-532
//                      do_FFT_Stockham_mk6<P>             ST/FFT-GPU-32bit-Stockham.cu:97-240
//                      do_FFT_Stockham_C2C<P,Dir>         RC/FFT-GPU-32bit-Stockham.cu:106-266
//                      do_FFT_Stockham_R2C_C2R<P,Dir>     RC/...:269-344
//   wrapper kernels    SMFFT_DIT_external / _multiple<P>  CT/...:534-572
//                      FFT_GPU_external / _multiple<P>    ST/...:243-278
//                      FFT_GPU_R2C_C2R_external / _multiple<P,Dir>   RC/...:349-384
// Same names, same template arguments, same calling contract:
//   * `s_input` is shared memory, natural (linear) order in and out, in place;
//   * blockDim.x == fft_length / 4 (four points per thread), 1-D block; a CT tile of fft_length
//     points holds fft_length / 2^fft_exp transforms (4 of 32, 2 of 64, else 1);
//   * the buffer needs no more than P::fft_sm_required elements (README.md:12) -- this
//     implementation uses exactly fft_length of them (XOR-swizzled exchanges need no padding);
//   * the caller synchronises before the call; do_SMFFT_CT_DIT does not synchronise on exit
//     (the caller does, CT:543-546), the Stockham functions end with __syncthreads() (ST:239).
// Internally (detail/compat_core.cuh): radix-4 register stages (4 points per thread IS the reference's contract).
// Tiles held by one warp (N <= 128) and every fft_reorder = 0 transform exchange through WARP SHUFFLES (detail/warp_fft.cuh:
// one 4x4 register/lane transposition per radix-4 stage, shared memory only across warps, the bit reversal folded into the
// addressing of the one store); natural-order transforms above 128 points run Stockham passes with exchanges in a
// bank-conflict-free swizzled layout inside the same buffer.  MUFU twiddles like the reference (no table pointer exists in this API).
//
// FFT_4096_inverse_noreorder::fft_direction is 1 here (mathematically correct).  Define
// SMFFT_COMPAT_QUIRK_4096 before including this header to reproduce the reference's 0
// (SM_FFT_parameters.cuh:388), which makes that instance run the forward transform.
#pragma once
#include "detail/compat_core.cuh"

// ---- Cooley-Tukey trait classes (member names and values as in SM_FFT_parameters.cuh) -------------
class FFT_Params {
public:
    static const int fft_exp = -1;
    static const int fft_length = -1;
    static const int warp = 32;
};

#define SMFFT_CT_PARAMS(NAME, EXP, SMREQ, LEN, DIR, REORDER)      \
    class NAME : public FFT_Params {                              \
    public:                                                       \
        static const int fft_exp = EXP;                           \
        static const int fft_sm_required = SMREQ;                 \
        static const int fft_length = LEN;                        \
        static const int fft_length_quarter = LEN / 4;            \
        static const int fft_length_half = LEN / 2;               \
        static const int fft_length_three_quarters = 3 * LEN / 4; \
        static const int fft_direction = DIR;                     \
        static const int fft_reorder = REORDER;                   \
    };
#define SMFFT_CT_PARAMS_4(N, EXP, SMREQ, LEN, INV_NOREORDER_DIR)              \
    SMFFT_CT_PARAMS(FFT_##N##_forward, EXP, SMREQ, LEN, 0, 1)                 \
    SMFFT_CT_PARAMS(FFT_##N##_forward_noreorder, EXP, SMREQ, LEN, 0, 0)       \
    SMFFT_CT_PARAMS(FFT_##N##_inverse, EXP, SMREQ, LEN, 1, 1)                 \
    SMFFT_CT_PARAMS(FFT_##N##_inverse_noreorder, EXP, SMREQ, LEN, INV_NOREORDER_DIR, 0)

#ifdef SMFFT_COMPAT_QUIRK_4096
#define SMFFT_4096_INV_NOREORDER_DIR 0
#else
#define SMFFT_4096_INV_NOREORDER_DIR 1
#endif

SMFFT_CT_PARAMS_4(32, 5, 128, 128, 1)      // four transforms per 128-point tile
SMFFT_CT_PARAMS_4(64, 6, 132, 128, 1)      // two transforms per tile
SMFFT_CT_PARAMS_4(128, 7, 132, 128, 1)
SMFFT_CT_PARAMS_4(256, 8, 264, 256, 1)     // (fft_length / 32) * 33 from here on (README.md:18)
SMFFT_CT_PARAMS_4(512, 9, 528, 512, 1)
SMFFT_CT_PARAMS_4(1024, 10, 1056, 1024, 1)
SMFFT_CT_PARAMS_4(2048, 11, 2112, 2048, 1)
SMFFT_CT_PARAMS_4(4096, 12, 4224, 4096, SMFFT_4096_INV_NOREORDER_DIR)
#undef SMFFT_CT_PARAMS_4
#undef SMFFT_CT_PARAMS

// ---- Stockham / R2C-C2R trait classes ------------------------------------------------------------
class FFT_ConstParams {
public:
    static const int fft_exp = -1;
    static const int fft_length = -1;
    static const int fft_half = -1;
    static const int warp = 32;
};
#define SMFFT_ST_PARAMS(N, EXP)                            \
    class FFT_##N : public FFT_ConstParams {               \
    public:                                                \
        static const int fft_exp = EXP;                    \
        static const int fft_quarter = N / 4;              \
        static const int fft_half = N / 2;                 \
        static const int fft_threequarters = 3 * N / 4;    \
        static const int fft_length = N;                   \
    };
SMFFT_ST_PARAMS(32, 5)   // 32..128: beyond the reference (its classes start at 256)
SMFFT_ST_PARAMS(64, 6)
SMFFT_ST_PARAMS(128, 7)
SMFFT_ST_PARAMS(256, 8)
SMFFT_ST_PARAMS(512, 9)
SMFFT_ST_PARAMS(1024, 10)
SMFFT_ST_PARAMS(2048, 11)
SMFFT_ST_PARAMS(4096, 12)
#undef SMFFT_ST_PARAMS

class FFT_ConstDirection {
public:
    static const int fft_direction = -1;
};
class FFT_forward : public FFT_ConstDirection {
public:
    static const int fft_direction = 0;
};
class FFT_inverse : public FFT_ConstDirection {
public:
    static const int fft_direction = 1;
};

// ---- device functions ------------------------------------------------------------------------------

template <class const_params>
__device__ __forceinline__ void do_SMFFT_CT_DIT(float2* s_input)
{
    smfft::compat::ct_dit<const_params::fft_exp, (const_params::fft_length >> const_params::fft_exp), const_params::fft_direction,
                          const_params::fft_reorder>(s_input);
}

template <class const_params, class const_direction>
__device__ __forceinline__ void do_FFT_Stockham_C2C(float2* s_input)
{
    using C = smfft::compat::Cfg<const_params::fft_exp, 1, const_direction::fft_direction, 1>;
    smfft::detail::block_fft_tile<C, smfft::detail::XF_C2C>(s_input, nullptr);
    __syncthreads();
}

// the reference's Stockham C2C directory is inverse-only (ST:70-78)
template <class const_params>
__device__ __forceinline__ void do_FFT_Stockham_mk6(float2* s_input)
{
    do_FFT_Stockham_C2C<const_params, FFT_inverse>(s_input);
}

// const_params::fft_length = N/2 complex points of a real transform of length N; packed bins
template <class const_params, class const_direction>
__device__ __forceinline__ void do_FFT_Stockham_R2C_C2R(float2* s_input)
{
    using C = smfft::compat::Cfg<const_params::fft_exp, 1, const_direction::fft_direction, 1>;
    if (const_direction::fft_direction == 0) {
        smfft::detail::block_fft_tile<C, smfft::detail::XF_C2C>(s_input, nullptr);
        __syncthreads();
        smfft::detail::r2c_pair_pass_tile<C, 0>(s_input, nullptr);
        __syncthreads();
    } else {
        smfft::detail::r2c_pair_pass_tile<C, 1>(s_input, nullptr);
        __syncthreads();
        smfft::detail::block_fft_tile<C, smfft::detail::XF_C2C>(s_input, nullptr);
        __syncthreads();
    }
}

// ---- wrapper kernels (launch contract of the reference: one CTA per fft_length tile) -----------------

#ifndef SMFFT_NREUSES
#define SMFFT_NREUSES 100  // NREUSES, CT/FFT-GPU-32bit.cu:10
#endif

namespace smfft {
namespace compat {

template <int LEN>
__device__ __forceinline__ void tile_in(float2* s, const float2* __restrict__ g)
{
    const size_t base = (size_t)blockIdx.x * LEN;  // 64-bit: the reference's 32-bit index stops at 4 GiB
#pragma unroll
    for (int q = 0; q < 4; q++) s[threadIdx.x + q * (LEN / 4)] = g[base + threadIdx.x + q * (LEN / 4)];
}
template <int LEN>
__device__ __forceinline__ void tile_out(const float2* s, float2* __restrict__ g)
{
    const size_t base = (size_t)blockIdx.x * LEN;
#pragma unroll
    for (int q = 0; q < 4; q++) g[base + threadIdx.x + q * (LEN / 4)] = s[threadIdx.x + q * (LEN / 4)];
}

}  // namespace compat
}  // namespace smfft

template <class const_params>
__global__ void __launch_bounds__(const_params::fft_length / 4) SMFFT_DIT_external(float2* d_input, float2* d_output)
{
    __shared__ float2 s_input[const_params::fft_sm_required];
    const size_t base = (size_t)blockIdx.x * const_params::fft_length;  // 64-bit: the reference's 32-bit index stops at 4 GiB
    smfft::compat::ct_dit_external<const_params::fft_exp, (const_params::fft_length >> const_params::fft_exp), const_params::fft_direction,
                                   const_params::fft_reorder>(s_input, d_input + base, d_output + base);
}

template <class const_params>
__global__ void __launch_bounds__(const_params::fft_length / 4) SMFFT_DIT_multiple(float2* d_input, float2* d_output)
{
    __shared__ __align__(16) float2 s_input[const_params::fft_sm_required];
    smfft::compat::tile_in<const_params::fft_length>(s_input, d_input);
    __syncthreads();
    for (int f = 0; f < SMFFT_NREUSES; f++) {
        do_SMFFT_CT_DIT<const_params>(s_input);
        __syncthreads();  // the reference omits this barrier, a cross-warp race for N >= 256
    }
    smfft::compat::tile_out<const_params::fft_length>(s_input, d_output);
}

template <class const_params>
__global__ void __launch_bounds__(const_params::fft_length / 4) FFT_GPU_external(float2* d_input, float2* d_output)
{
    extern __shared__ float2 s_input_dyn[];  // FFT_size * 8 bytes (ST:319)
    smfft::compat::tile_in<const_params::fft_length>(s_input_dyn, d_input);
    __syncthreads();
    do_FFT_Stockham_mk6<const_params>(s_input_dyn);
    smfft::compat::tile_out<const_params::fft_length>(s_input_dyn, d_output);
}

template <class const_params>
__global__ void __launch_bounds__(const_params::fft_length / 4) FFT_GPU_multiple(float2* d_input, float2* d_output)
{
    extern __shared__ float2 s_input_dyn[];
    smfft::compat::tile_in<const_params::fft_length>(s_input_dyn, d_input);
    __syncthreads();
    for (int f = 0; f < 100; f++) do_FFT_Stockham_mk6<const_params>(s_input_dyn);
    smfft::compat::tile_out<const_params::fft_length>(s_input_dyn, d_output);
}

template <class const_params, class const_direction>
__global__ void __launch_bounds__(const_params::fft_length / 4) FFT_GPU_R2C_C2R_external(float2* d_input, float2* d_output)
{
    __shared__ float2 s_input[const_params::fft_length + 1];
    smfft::compat::tile_in<const_params::fft_length>(s_input, d_input);
    __syncthreads();
    do_FFT_Stockham_R2C_C2R<const_params, const_direction>(s_input);
    smfft::compat::tile_out<const_params::fft_length>(s_input, d_output);
}

template <class const_params, class const_direction>
__global__ void __launch_bounds__(const_params::fft_length / 4) FFT_GPU_R2C_C2R_multiple(float2* d_input, float2* d_output)
{
    __shared__ float2 s_input[const_params::fft_length + 1];
    smfft::compat::tile_in<const_params::fft_length>(s_input, d_input);
    __syncthreads();
    for (int f = 0; f < 100; f++) do_FFT_Stockham_R2C_C2R<const_params, const_direction>(s_input);
    smfft::compat::tile_out<const_params::fft_length>(s_input, d_output);
}
