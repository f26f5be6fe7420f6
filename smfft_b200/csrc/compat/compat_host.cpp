// compat_host.cpp -- libsmfft_compat.so: the reference's host symbols forwarded onto the C ABI.
// See include/smfft_compat.hpp for the reference lines each function replaces.
#include <cufft.h>
#include <stdio.h>
#include <stdlib.h>

#include "smfft.h"
#include "smfft_compat.hpp"

#define VIS __attribute__((visibility("default")))

static void report(int rc, const char* what)
{
    if (rc) fprintf(stderr, "smfft (%s): %s\n", what, smfft_last_error());
}

// The reference tells two failures of a launcher apart: an unsupported FFT length only prints "Error wrong FFT length!"
// and the launcher carries on (CT:656-658, ST:337-339, RC:425-427); anything CUDA goes through checkCudaErrors, which
// prints the error and ends the process (utils_cuda.h:12-22).  Same here, from the C ABI's error code.
static void launcher_failed(const char* what)
{
    if (smfft_last_error_code() == SMFFT_ERR_ARGUMENT) {
        printf("Error wrong FFT length!\n");
        return;
    }
    fprintf(stderr, "CUDA error in %s: %s\n", what, smfft_last_error());
    exit(1);
}

VIS void FFT_init() { report(smfft_init(), "FFT_init"); }

// ---- Cooley-Tukey ----
VIS int FFT_external_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, bool inverse, bool reorder, double* FFT_time)
{
    if (smfft_external_benchmark(d_input, d_output, FFT_size, nFFTs, inverse, reorder, FFT_time)) launcher_failed("FFT_external_benchmark");
    return 0;  // CT:656-664: the reference returns 0 after the wrong-length message too
}
VIS int FFT_multiple_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, bool inverse, bool reorder, double* FFT_time)
{
    if (nFFTs / 100 == 0) {  // CT:669-673
        *FFT_time = -1;
        return 1;
    }
    if (smfft_multiple_benchmark(d_input, d_output, FFT_size, nFFTs, inverse, reorder, FFT_time)) launcher_failed("FFT_multiple_benchmark");
    return 0;
}
VIS int GPU_smFFT_4elements(float2* h_input, float2* h_output, int FFT_size, int nFFTs, bool inverse, bool reorder, int nRuns,
                            double* single_ex_time, double* multi_ex_time)
{
    if (FFT_size == 32 && (nFFTs % 4) != 0) return 1;  // CT:835-836
    if (FFT_size == 64 && (nFFTs % 2) != 0) return 1;
    double s = 0, m = 0;
    int rc = smfft_c2c_host(h_input, h_output, FFT_size, nFFTs, inverse, reorder, nRuns, &s, &m);
    report(rc, "GPU_smFFT_4elements");
    if (rc) return 1;
    *single_ex_time = s;
    *multi_ex_time = m;
    printf("  SH FFT normal = %0.3f ms; SM FFT multiple times = %0.3f ms\n", s, m);  // CT:895
    return 0;
}

static int cufft_c2c_host(float2* h_in, float2* h_out, int n, int nffts, int dir, double* ms)
{
    const size_t bytes = (size_t)n * nffts * sizeof(float2);
    float2 *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_in, bytes) != cudaSuccess || cudaMalloc(&d_out, bytes) != cudaSuccess) {
        printf("Error: Not enough memory! Input data are too big for the device.\n");
        cudaFree(d_in);
        return 1;
    }
    cudaMemcpy(d_in, h_in, bytes, cudaMemcpyHostToDevice);
    cufftHandle plan;
    cufftResult e = cufftPlan1d(&plan, n, CUFFT_C2C, nffts);
    if (e != CUFFT_SUCCESS) printf("CUFFT error: %d", e);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, 0);
    cufftExecC2C(plan, (cufftComplex*)d_in, (cufftComplex*)d_out, dir);
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    cufftDestroy(plan);
    *ms = t;
    printf("  FFT size: %d; cuFFT time = %0.3f ms;\n", n, t);
    cudaMemcpy(h_out, d_out, bytes, cudaMemcpyDeviceToHost);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_in);
    cudaFree(d_out);
    return 0;
}
VIS int GPU_cuFFT(float2* h_input, float2* h_output, int FFT_size, int nFFTs, bool inverse, int, double* single_ex_time)
{
    return cufft_c2c_host(h_input, h_output, FFT_size, nFFTs, inverse ? CUFFT_INVERSE : CUFFT_FORWARD, single_ex_time);
}

// ---- Stockham C2C ----
VIS void FFT_external_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, double* FFT_time)
{
    if (smfft_stockham_external_benchmark(d_input, d_output, FFT_size, nFFTs, 1, FFT_time)) launcher_failed("FFT_external_benchmark");
}
VIS void FFT_multiple_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, double* FFT_time)
{
    if (nFFTs / 100 == 0) return;
    if (smfft_stockham_multiple_benchmark(d_input, d_output, FFT_size, nFFTs, 1, FFT_time)) launcher_failed("FFT_multiple_benchmark");
}
VIS int GPU_FFT_C2C_Stockham(float2* h_input, float2* h_output, int FFT_size, int nFFTs, int nRuns, double* single_ex_time,
                             double* multi_ex_time)
{
    double s = 0, m = 0;
    int rc = smfft_c2c_host(h_input, h_output, FFT_size, nFFTs, 1, 1, nRuns, &s, &m);
    report(rc, "GPU_FFT_C2C_Stockham");
    if (rc) return 1;
    *single_ex_time = s;
    *multi_ex_time = m;
    printf("  SH FFT normal = %0.3f ms; SM FFT multiple times = %0.3f ms\n", s, m);
    return 0;
}
VIS int GPU_cuFFT(float2* h_input, float2* h_output, int FFT_size, int nFFTs, int, double* single_ex_time)
{
    return cufft_c2c_host(h_input, h_output, FFT_size, nFFTs, CUFFT_INVERSE, single_ex_time);  // ST:429
}

// ---- Stockham R2C / C2R ----
VIS void FFT_external_benchmark(float* d_input, float* d_output, int FFT_size, int nFFTs, int inverse, double* FFT_time)
{
    if (smfft_r2c_c2r_external_benchmark(d_input, d_output, FFT_size, nFFTs, inverse, FFT_time)) launcher_failed("FFT_external_benchmark");
}
VIS void FFT_multiple_benchmark(float* d_input, float* d_output, int FFT_size, int nFFTs, double* FFT_time)
{
    if (nFFTs / 100 == 0) return;
    if (smfft_r2c_multiple_benchmark(d_input, d_output, FFT_size, nFFTs, FFT_time)) launcher_failed("FFT_multiple_benchmark");
}
VIS int GPU_smFFT_R2C(float2* h_output, float* h_input, int FFT_size, int nFFTs, int nRuns)
{
    double s = 0, m = 0;
    int rc = smfft_r2c_c2r_host(h_input, h_output, FFT_size, nFFTs, 0, nRuns, &s, &m);
    report(rc, "GPU_smFFT_R2C");
    if (!rc) printf("  SH FFT normal = %0.3f ms; SM FFT multiple times = %0.3f ms\n", s, m);
    return rc ? 1 : 0;
}
VIS int GPU_smFFT_C2R(float* h_output, float2* h_input, int FFT_size, int nFFTs, int nRuns)
{
    double s = 0, m = 0;
    int rc = smfft_r2c_c2r_host(h_input, h_output, FFT_size, nFFTs, 1, nRuns, &s, nullptr);
    report(rc, "GPU_smFFT_C2R");
    if (!rc) printf("  SH FFT normal = %0.3f ms; SM FFT multiple times = %0.3f ms\n", s, m);
    return rc ? 1 : 0;
}

static int cufft_real_host(void* h_out, void* h_in, int n, int nffts, int nRuns, bool c2r)
{
    const size_t rbytes = (size_t)n * nffts * sizeof(float), cbytes = (size_t)(n / 2 + 1) * nffts * sizeof(float2);
    void *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_in, c2r ? cbytes : rbytes) != cudaSuccess || cudaMalloc(&d_out, c2r ? rbytes : cbytes) != cudaSuccess) {
        printf("Error: Not enough memory!\n");
        cudaFree(d_in);
        return 1;
    }
    cudaMemcpy(d_in, h_in, c2r ? cbytes : rbytes, cudaMemcpyHostToDevice);
    cufftHandle plan;
    cufftResult e = cufftPlan1d(&plan, n, c2r ? CUFFT_C2R : CUFFT_R2C, nffts);
    if (e != CUFFT_SUCCESS) printf("CUFFT error: %d", e);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, 0);
    for (int r = 0; r < (nRuns < 1 ? 1 : nRuns); r++) {  // RC:502-504 loops nRuns inside the timer
        if (c2r) cufftExecC2R(plan, (cufftComplex*)d_in, (cufftReal*)d_out);
        else cufftExecR2C(plan, (cufftReal*)d_in, (cufftComplex*)d_out);
    }
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    cufftDestroy(plan);
    printf("  FFT size: %d; cuFFT %s time = %0.3f ms;\n", n, c2r ? "C2R" : "R2C", t / (nRuns < 1 ? 1 : nRuns));
    cudaMemcpy(h_out, d_out, c2r ? rbytes : cbytes, cudaMemcpyDeviceToHost);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_in);
    cudaFree(d_out);
    return 0;
}
VIS int GPU_cuFFT_R2C(float2* h_output, float* h_input, int FFT_size, int nFFTs, int nRuns)
{
    return cufft_real_host(h_output, h_input, FFT_size, nFFTs, nRuns, false);
}
VIS int GPU_cuFFT_C2R(float* h_output, float2* h_input, int FFT_size, int nFFTs, int nRuns)
{
    return cufft_real_host(h_output, h_input, FFT_size, nFFTs, nRuns, true);
}
