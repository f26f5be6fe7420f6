// tools/fftlike_copy.cu -- does a plain copy with cuFFT's access pattern reach cuFFT's 1.22 ms?  (measurement tool)
//
// cuFFT's 1024-point kernel (vector_fft<1024, EPT<32>, 4, ...>, profiles/r01_ncu_summary.md): 128 threads, 4 transforms
// per CTA, one CTA per 32 KB, every thread loads 32 x 8 bytes up front (warp w streams through its own 8 KB), computes,
// stores 32 x 8 bytes.  This program times copies with that shape: PATTERN 0 = warp-private 8 KB streams, 1 = CTA-wide
// 1 KB rows; DELAY = cycles spent between the loads and the stores (stands in for the FFT); CTAs per SM limited by a
// dynamic shared-memory pad.  Output: CSV.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

template <int PATTERN, int CACHE>
__global__ void __launch_bounds__(128) fftlike(const float2* __restrict__ a, float2* __restrict__ b, int delay)
{
    extern __shared__ unsigned char pad[];
    const size_t base = (size_t)blockIdx.x * 4096;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    float2 v[32];
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const size_t i = PATTERN == 0 ? base + w * 1024 + k * 32 + l : base + k * 128 + threadIdx.x;
        if (CACHE == 0) v[k] = a[i];
        else asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[k].x), "=f"(v[k].y) : "l"(a + i));
    }
    if (delay > 0) {
        // wait for the data, then burn `delay` cycles (the stores must not start before)
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k++) s += v[k].x;
        const long long t0 = clock64();
        while (clock64() - t0 < delay) {}
        if (s == 123.456f) v[0].y += 1.f;
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const size_t i = PATTERN == 0 ? base + w * 1024 + k * 32 + l : base + k * 128 + threadIdx.x;
        if (CACHE == 0) b[i] = v[k];
        else asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(b + i), "f"(v[k].x), "f"(v[k].y) : "memory");
    }
}

int main(int argc, char** argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 29;
    const size_t pts = (size_t)1 << lg;
    float2 *a, *b;
    CK(cudaMalloc(&a, pts * 8));
    CK(cudaMalloc(&b, pts * 8));
    CK(cudaMemset(a, 1, pts * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("pattern,cache,delay,ctas_per_sm,ms_med,ms_min,GBps\n");
    auto run = [&](auto kern, int pattern, int cache, int delay, int per) {
        const size_t smem = (size_t)((220 * 1024 / per) & ~1023);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        std::vector<float> t;
        for (int r = 0; r < 9; r++) {
            CK(cudaEventRecord(e0));
            kern<<<(unsigned)(pts / 4096), 128, smem>>>(a, b, delay);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        printf("%d,%d,%d,%d,%.4f,%.4f,%.1f\n", pattern, cache, delay, per, t[t.size() / 2], t[0], pts * 16.0 / t[t.size() / 2] / 1e6);
        fflush(stdout);
    };
    for (int per : {3, 4, 6, 8})
        for (int delay : {0, 1500, 3000, 6000}) {
            run(fftlike<0, 0>, 0, 0, delay, per);
            run(fftlike<1, 0>, 1, 0, delay, per);
            run(fftlike<0, 1>, 0, 1, delay, per);
            run(fftlike<1, 1>, 1, 1, delay, per);
        }
    return 0;
}
