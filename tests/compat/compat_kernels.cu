// compat_kernels.cu -- a "user program" for the drop-in device API: it includes smfft/compat.cuh
// and launches the wrapper kernels by the reference's own names, template arguments and launch
// shapes (CT/FFT-GPU-32bit.cu:586-659, ST/...:309-335, RC/...:399-425).  Built by the tests into
// tests/compat/_build/libcompat_kernels.so and driven through ctypes on the GPU.
#include <cuda_runtime.h>

#include "smfft/compat.cuh"
#include "smfft/device.cuh"

#define CT_CASE(N, KERNEL)                                                                                     \
    case N:                                                                                                    \
        if (!inverse && reorder) KERNEL<FFT_##N##_forward><<<grid, block>>>(d_in, d_out);                      \
        if (!inverse && !reorder) KERNEL<FFT_##N##_forward_noreorder><<<grid, block>>>(d_in, d_out);           \
        if (inverse && reorder) KERNEL<FFT_##N##_inverse><<<grid, block>>>(d_in, d_out);                       \
        if (inverse && !reorder) KERNEL<FFT_##N##_inverse_noreorder><<<grid, block>>>(d_in, d_out);            \
        break;

static void ct_shape(int n, int nffts, int reuse, dim3* grid, dim3* block)
{
    *grid = dim3(nffts / reuse);
    *block = dim3(n / 4);
    if (n == 32) { *grid = dim3(nffts / (4 * reuse)); *block = dim3(32); }
    if (n == 64) { *grid = dim3(nffts / (2 * reuse)); *block = dim3(32); }
}

extern "C" int compat_ct_external(float2* d_in, float2* d_out, int n, int nffts, int inverse, int reorder)
{
    dim3 grid, block;
    ct_shape(n, nffts, 1, &grid, &block);
    switch (n) {
        CT_CASE(32, SMFFT_DIT_external) CT_CASE(64, SMFFT_DIT_external) CT_CASE(128, SMFFT_DIT_external)
        CT_CASE(256, SMFFT_DIT_external) CT_CASE(512, SMFFT_DIT_external) CT_CASE(1024, SMFFT_DIT_external)
        CT_CASE(2048, SMFFT_DIT_external) CT_CASE(4096, SMFFT_DIT_external)
        default: return -1;
    }
    return (int)cudaGetLastError();
}

extern "C" int compat_ct_multiple(float2* d_in, float2* d_out, int n, int nffts, int inverse, int reorder)
{
    dim3 grid, block;
    ct_shape(n, nffts, 100, &grid, &block);
    if (grid.x == 0) return -2;
    switch (n) {
        CT_CASE(32, SMFFT_DIT_multiple) CT_CASE(64, SMFFT_DIT_multiple) CT_CASE(128, SMFFT_DIT_multiple)
        CT_CASE(256, SMFFT_DIT_multiple) CT_CASE(512, SMFFT_DIT_multiple) CT_CASE(1024, SMFFT_DIT_multiple)
        CT_CASE(2048, SMFFT_DIT_multiple) CT_CASE(4096, SMFFT_DIT_multiple)
        default: return -1;
    }
    return (int)cudaGetLastError();
}

extern "C" int compat_stockham_external(float2* d_in, float2* d_out, int n, int nffts)
{
    switch (n) {
        case 256: FFT_GPU_external<FFT_256><<<nffts, n / 4, n * 8>>>(d_in, d_out); break;
        case 512: FFT_GPU_external<FFT_512><<<nffts, n / 4, n * 8>>>(d_in, d_out); break;
        case 1024: FFT_GPU_external<FFT_1024><<<nffts, n / 4, n * 8>>>(d_in, d_out); break;
        case 2048: FFT_GPU_external<FFT_2048><<<nffts, n / 4, n * 8>>>(d_in, d_out); break;
        case 4096: FFT_GPU_external<FFT_4096><<<nffts, n / 4, n * 8>>>(d_in, d_out); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}

extern "C" int compat_r2c_c2r_external(float* d_in, float* d_out, int n, int nffts, int inverse)
{
    float2* in = (float2*)d_in;
    float2* out = (float2*)d_out;
    const int block = (n >> 1) / 4;
#define RC_CASE(N, CORE)                                                                          \
    case N:                                                                                       \
        if (!inverse) FFT_GPU_R2C_C2R_external<FFT_##CORE, FFT_forward><<<nffts, block>>>(in, out); \
        else FFT_GPU_R2C_C2R_external<FFT_##CORE, FFT_inverse><<<nffts, block>>>(in, out);          \
        break;
    switch (n) {
        RC_CASE(256, 128) RC_CASE(512, 256) RC_CASE(1024, 512) RC_CASE(2048, 1024) RC_CASE(4096, 2048)
        default: return -1;
    }
    return (int)cudaGetLastError();
}

extern "C" int compat_stockham_multiple(float2* d_in, float2* d_out, int n, int nffts)
{
    const int grid = nffts / 100;  // ST:351
    if (grid == 0) return -2;
    switch (n) {
        case 256: FFT_GPU_multiple<FFT_256><<<grid, n / 4, n * 8>>>(d_in, d_out); break;
        case 512: FFT_GPU_multiple<FFT_512><<<grid, n / 4, n * 8>>>(d_in, d_out); break;
        case 1024: FFT_GPU_multiple<FFT_1024><<<grid, n / 4, n * 8>>>(d_in, d_out); break;
        case 2048: FFT_GPU_multiple<FFT_2048><<<grid, n / 4, n * 8>>>(d_in, d_out); break;
        case 4096: FFT_GPU_multiple<FFT_4096><<<grid, n / 4, n * 8>>>(d_in, d_out); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}

extern "C" int compat_r2c_multiple(float* d_in, float* d_out, int n, int nffts)
{
    float2* in = (float2*)d_in;
    float2* out = (float2*)d_out;
    const int grid = nffts / 100, block = (n >> 1) / 4;  // RC:438-439
    if (grid == 0) return -2;
    switch (n) {
        case 256: FFT_GPU_R2C_C2R_multiple<FFT_128, FFT_forward><<<grid, block>>>(in, out); break;
        case 512: FFT_GPU_R2C_C2R_multiple<FFT_256, FFT_forward><<<grid, block>>>(in, out); break;
        case 1024: FFT_GPU_R2C_C2R_multiple<FFT_512, FFT_forward><<<grid, block>>>(in, out); break;
        case 2048: FFT_GPU_R2C_C2R_multiple<FFT_1024, FFT_forward><<<grid, block>>>(in, out); break;
        case 4096: FFT_GPU_R2C_C2R_multiple<FFT_2048, FFT_forward><<<grid, block>>>(in, out); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}

// a user kernel that calls the device function directly: forward (no-reorder) -> pointwise filter
// -> inverse in one launch, the convolution use case SMFFT exists for (README.md:2, 10-14).
// Filter H is given in natural frequency order; with the no-reorder pair the data is transformed
// by F*P, so the filter is applied in the same order: y = (1/N) F^-1( H .* (F x) ) needs reorder=1
// transforms; this kernel uses the reorder pair for clarity and checks against numpy in the test.
template <class FWD, class INV>
__global__ void user_convolve(const float2* x, const float2* H, float2* y)
{
    __shared__ float2 s[FWD::fft_sm_required];
    smfft::compat::tile_in<FWD::fft_length>(s, x);
    __syncthreads();
    do_SMFFT_CT_DIT<FWD>(s);
    __syncthreads();
    for (int q = 0; q < 4; q++) {
        const int i = threadIdx.x + q * (FWD::fft_length / 4);
        const float2 a = s[i], h = H[i];
        s[i] = make_float2((a.x * h.x - a.y * h.y) / FWD::fft_length, (a.x * h.y + a.y * h.x) / FWD::fft_length);
    }
    __syncthreads();
    do_SMFFT_CT_DIT<INV>(s);
    __syncthreads();
    smfft::compat::tile_out<FWD::fft_length>(s, y);
}

extern "C" int compat_user_convolve_1024(const float2* x, const float2* H, float2* y, int nffts)
{
    user_convolve<FFT_1024_forward, FFT_1024_inverse><<<nffts, 256>>>(x, H, y);
    return (int)cudaGetLastError();
}

// ---- the NATIVE device primitive (include/smfft/device.cuh): 16 points per thread, data in registers across the call ----

template <class FFT>
__global__ void __launch_bounds__(FFT::THREADS) native_fft(const float2* x, float2* y, const float2* w8192)
{
    __shared__ __align__(16) float2 xch[FFT::EXCHANGE_POINTS];
    __shared__ float2 tw[FFT::TWIDDLE_POINTS + 1];
    if (FFT::TWIDDLE_POINTS) {
        FFT::fill_twiddles(tw, w8192);
        __syncthreads();
    }
    float2 v[FFT::R];
    const size_t base = (size_t)blockIdx.x * FFT::TILE_POINTS;
    FFT::load(v, x + base);
    FFT::exec(v, xch, tw);
    FFT::store(v, y + base);
}

// lut != 0: table twiddles (w8192 = smfft_twiddle_table()), else MUFU; r32 != 0: 32 points per thread
extern "C" int native_fft_launch(const float2* x, float2* y, int n, int nffts, int inverse, int lut, const float2* w8192)
{
#define NF(E, F)                                                                                                                    \
    case (1 << E): {                                                                                                                \
        const int grid = nffts / F;                                                                                                 \
        if (!inverse && !lut) native_fft<smfft::BlockFFT<E, smfft::FORWARD, F>><<<grid, smfft::BlockFFT<E, 0, F>::THREADS>>>(x, y, w8192);              \
        if (inverse && !lut) native_fft<smfft::BlockFFT<E, smfft::INVERSE, F>><<<grid, smfft::BlockFFT<E, 0, F>::THREADS>>>(x, y, w8192);               \
        if (!inverse && lut) native_fft<smfft::BlockFFT<E, smfft::FORWARD, F, smfft::TW_LUT>><<<grid, smfft::BlockFFT<E, 0, F>::THREADS>>>(x, y, w8192); \
        if (inverse && lut) native_fft<smfft::BlockFFT<E, smfft::INVERSE, F, smfft::TW_LUT>><<<grid, smfft::BlockFFT<E, 0, F>::THREADS>>>(x, y, w8192);  \
        break;                                                                                                                      \
    }
    switch (n) {
        NF(5, 32) NF(6, 16) NF(7, 8) NF(8, 4) NF(9, 2) NF(10, 2) NF(11, 1) NF(12, 1)
        default: return -1;
    }
#undef NF
    return (int)cudaGetLastError();
}

// the convolution use case on the native primitive: load -> forward -> multiply by H (in registers) -> inverse -> store
template <int E, int F, int TW>
__global__ void __launch_bounds__(smfft::BlockFFT<E, 0, F>::THREADS) native_convolve(const float2* x, const float2* H, float2* y, const float2* w8192)
{
    using FWD = smfft::BlockFFT<E, smfft::FORWARD, F, TW>;
    using INV = smfft::BlockFFT<E, smfft::INVERSE, F, TW>;
    __shared__ __align__(16) float2 xch[FWD::EXCHANGE_POINTS];
    __shared__ float2 twf[FWD::TWIDDLE_POINTS + 1], twi[FWD::TWIDDLE_POINTS + 1];
    if (FWD::TWIDDLE_POINTS) {
        FWD::fill_twiddles(twf, w8192);
        INV::fill_twiddles(twi, w8192);
        __syncthreads();
    }
    float2 v[FWD::R];
    const size_t base = (size_t)blockIdx.x * FWD::TILE_POINTS;
    FWD::load(v, x + base);
    smfft::block_convolve<FWD, INV>(v, xch, [&](float2 a, int k) {
        const float2 h = __ldg(H + k);
        constexpr float sc = 1.0f / (float)FWD::N;
        return make_float2((a.x * h.x - a.y * h.y) * sc, (a.x * h.y + a.y * h.x) * sc);
    }, twf, twi);
    INV::store(v, y + base);
}

extern "C" int native_convolve_launch(const float2* x, const float2* H, float2* y, int n, int nffts, int lut, const float2* w8192)
{
    switch (n) {
        case 256: if (lut) native_convolve<8, 4, smfft::TW_LUT><<<nffts / 4, 64>>>(x, H, y, w8192); else native_convolve<8, 4, smfft::TW_MUFU><<<nffts / 4, 64>>>(x, H, y, w8192); break;
        case 1024: if (lut) native_convolve<10, 2, smfft::TW_LUT><<<nffts / 2, 128>>>(x, H, y, w8192); else native_convolve<10, 2, smfft::TW_MUFU><<<nffts / 2, 128>>>(x, H, y, w8192); break;
        case 4096: if (lut) native_convolve<12, 1, smfft::TW_LUT><<<nffts, 256>>>(x, H, y, w8192); else native_convolve<12, 1, smfft::TW_MUFU><<<nffts, 256>>>(x, H, y, w8192); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}

extern "C" int compat_user_convolve(const float2* x, const float2* H, float2* y, int n, int nffts)
{
    switch (n) {
        case 256: user_convolve<FFT_256_forward, FFT_256_inverse><<<nffts, 64>>>(x, H, y); break;
        case 1024: user_convolve<FFT_1024_forward, FFT_1024_inverse><<<nffts, 256>>>(x, H, y); break;
        case 4096: user_convolve<FFT_4096_forward, FFT_4096_inverse><<<nffts, 1024>>>(x, H, y); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}
