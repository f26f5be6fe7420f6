// smfft_compat.hpp -- the reference's HOST interface, C++ linkage, exact signatures.
//
// KAdamek/SMFFT has no FFI: each of its three FFT.c programs (compiled as C++, CT/Makefile:22)
// links mangled C++ launchers out of the matching .cu file.  libsmfft_compat.so defines every one
// of those symbols as a thin forwarder onto the C ABI (include/smfft.h), so the reference's FFT.c
// objects link and run unmodified against this library (recipe: oracle/Makefile target `refmain`;
// INTEGRATION.md).  The three programs' overloads differ in parameter lists, so they coexist.
//
//   SMFFT_CooleyTukey_C2C/FFT.c:80-81, FFT-GPU-32bit.cu:576, 583, 666, 758, 827
//   SMFFT_Stockham_C2C/FFT.c:79-81,    FFT-GPU-32bit-Stockham.cu:299, 306, 348, 389, 457
//   SMFFT_Stockham_R2C_C2R/FFT.c:188-191, FFT-GPU-32bit-Stockham.cu:388, 396, 435, 471, 520, 572, 631
//
// GPU_cuFFT* are the reference's comparison baselines (they call cuFFT by design); they live only
// in the compat library, never in libsmfft.so.
#pragma once
#include <cuda_runtime.h>

void FFT_init();

// ---- Cooley-Tukey C2C ----
int FFT_external_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, bool inverse, bool reorder, double* FFT_time);
int FFT_multiple_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, bool inverse, bool reorder, double* FFT_time);
int GPU_smFFT_4elements(float2* h_input, float2* h_output, int FFT_size, int nFFTs, bool inverse, bool reorder, int nRuns,
                        double* single_ex_time, double* multi_ex_time);
int GPU_cuFFT(float2* h_input, float2* h_output, int FFT_size, int nFFTs, bool inverse, int nRuns, double* single_ex_time);

// ---- Stockham C2C (inverse only in the reference) ----
void FFT_external_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, double* FFT_time);
void FFT_multiple_benchmark(float2* d_input, float2* d_output, int FFT_size, int nFFTs, double* FFT_time);
int GPU_FFT_C2C_Stockham(float2* h_input, float2* h_output, int FFT_size, int nFFTs, int nRuns, double* single_ex_time,
                         double* multi_ex_time);
int GPU_cuFFT(float2* h_input, float2* h_output, int FFT_size, int nFFTs, int nRuns, double* single_ex_time);

// ---- Stockham R2C / C2R ----
void FFT_external_benchmark(float* d_input, float* d_output, int FFT_size, int nFFTs, int inverse, double* FFT_time);
void FFT_multiple_benchmark(float* d_input, float* d_output, int FFT_size, int nFFTs, double* FFT_time);
int GPU_smFFT_R2C(float2* h_output, float* h_input, int FFT_size, int nFFTs, int nRuns);
int GPU_smFFT_C2R(float* h_output, float2* h_input, int FFT_size, int nFFTs, int nRuns);
int GPU_cuFFT_R2C(float2* h_output, float* h_input, int FFT_size, int nFFTs, int nRuns);
int GPU_cuFFT_C2R(float* h_output, float2* h_input, int FFT_size, int nFFTs, int nRuns);
