// oracle/ref_user_convolve.cu -- TEST / BENCH INFRASTRUCTURE, never part of the product.
// A user kernel around the REFERENCE's own device function: load -> do_SMFFT_CT_DIT<forward> -> pointwise multiply ->
// do_SMFFT_CT_DIT<inverse> -> store (the use case README.md:2, 10-14 of the reference describes).  The reference's
// translation unit is included from where it lies (-I$(REF)/SMFFT_CooleyTukey_C2C, oracle/Makefile target `ref`);
// nothing of it is copied here.  Built into oracle/_ref/libsmfft_ref_conv.so; tools/convolve_bench.py times it next to the
// same kernel on include/smfft/compat.cuh and on the native primitive include/smfft/device.cuh.
#include "FFT-GPU-32bit.cu"

template <class FWD, class INV>
__global__ void ref_user_convolve(const float2* x, const float2* H, float2* y)
{
    __shared__ float2 s[FWD::fft_sm_required];
    const size_t base = (size_t)blockIdx.x * FWD::fft_length;
    for (int q = 0; q < 4; q++) s[threadIdx.x + q * (FWD::fft_length / 4)] = x[base + threadIdx.x + q * (FWD::fft_length / 4)];
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
    for (int q = 0; q < 4; q++) y[base + threadIdx.x + q * (FWD::fft_length / 4)] = s[threadIdx.x + q * (FWD::fft_length / 4)];
}

extern "C" int ref_user_convolve_launch(const float2* x, const float2* H, float2* y, int n, int nffts)
{
    switch (n) {
        case 256: ref_user_convolve<FFT_256_forward, FFT_256_inverse><<<nffts, 64>>>(x, H, y); break;
        case 1024: ref_user_convolve<FFT_1024_forward, FFT_1024_inverse><<<nffts, 256>>>(x, H, y); break;
        case 4096: ref_user_convolve<FFT_4096_forward, FFT_4096_inverse><<<nffts, 1024>>>(x, H, y); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}
