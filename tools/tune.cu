// tools/tune.cu -- shape sweep of the native kernel on one GPU (measurement tool, not product).
//
//   nvcc ... tools/tune.cu -o tools/tune && tools/tune [log2_points=29] [reps=5] [only_e=0] [real=0] > gpurun_out/tune.csv
//
// For every FFT size it times a list of kernel shapes (points/thread, tile size, pipeline stages,
// CTAs/SM, TMA vs LDG staging, table vs MUFU twiddles) on a 4 GiB batch with CUDA events, checks
// each variant's output against the first variant of that size, and prints one CSV row per variant.
// The winners go into smfft_b200/csrc/tuning.hpp; the CSV is kept under profiles/.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "tmap.hpp"
#include "tune_shapes.cuh"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));  \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

std::vector<Variant> g_variants;
void add_sizes_a();
void add_sizes_b();
void add_sizes_c();
void add_sizes_d();
void add_sizes_real_a();
void add_sizes_real_b();
void add_sizes_real_c();
void add_sizes_reg_a();
void add_sizes_reg_b();
void add_sizes_reg_c();
void add_sizes_reg_d();
void add_sizes_reg_e();
void add_sizes_dual_a();
void add_sizes_dual_b();
void add_sizes_dual_c();
void add_sizes_dual_d();
void add_sizes_dual_e();
#ifdef TUNE_ONLY_REG_C  // cache-operator experiments: only the 1024-point register-direct shapes are linked
void add_sizes_a() {}
void add_sizes_b() {}
void add_sizes_c() {}
void add_sizes_d() {}
void add_sizes_real_a() {}
void add_sizes_real_b() {}
void add_sizes_real_c() {}
void add_sizes_reg_a() {}
void add_sizes_reg_b() {}
void add_sizes_reg_d() {}
void add_sizes_reg_e() {}
#endif
#ifdef TUNE_ONLY_REG  // only the register-direct sweep is linked
void add_sizes_a() {}
void add_sizes_b() {}
void add_sizes_c() {}
void add_sizes_d() {}
void add_sizes_real_a() {}
void add_sizes_real_b() {}
void add_sizes_real_c() {}
void add_sizes_dual_a() {}
void add_sizes_dual_b() {}
void add_sizes_dual_c() {}
void add_sizes_dual_d() {}
void add_sizes_dual_e() {}
#endif
#ifdef TUNE_ONLY_DUAL  // only the dual-lane sweep is linked
void add_sizes_a() {}
void add_sizes_b() {}
void add_sizes_c() {}
void add_sizes_d() {}
void add_sizes_real_a() {}
void add_sizes_real_b() {}
void add_sizes_real_c() {}
void add_sizes_reg_a() {}
void add_sizes_reg_b() {}
void add_sizes_reg_c() {}
void add_sizes_reg_d() {}
void add_sizes_reg_e() {}
#endif
static void add_all_sizes(int real)
{
    if (real == 3) {
        add_sizes_dual_a(); add_sizes_dual_b(); add_sizes_dual_c(); add_sizes_dual_d(); add_sizes_dual_e();
    } else if (real == 2) {
        add_sizes_reg_a(); add_sizes_reg_b(); add_sizes_reg_c(); add_sizes_reg_d(); add_sizes_reg_e();
    } else if (real) {
        add_sizes_real_a(); add_sizes_real_b(); add_sizes_real_c();
    } else {
        add_sizes_a(); add_sizes_b(); add_sizes_c(); add_sizes_d();
    }
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 v0 = a[i], v1 = a[i + stride], v2 = a[i + 2 * stride], v3 = a[i + 3 * stride];
        b[i] = v0; b[i + stride] = v1; b[i + 2 * stride] = v2; b[i + 3 * stride] = v3;
    }
    for (; i < n; i += stride) b[i] = a[i];
}

__global__ void fill_kernel(float* p, size_t n, unsigned seed)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        p[i] = (float)(z >> 40) * (1.0f / 16777216.0f);
    }
}

int main(int argc, char** argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 29;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const int only_e = argc > 3 ? atoi(argv[3]) : 0;
    const int real = argc > 4 ? atoi(argv[4]) : 0;  // 1: sweep the R2C / C2R shapes instead of C2C, 2: register-direct input shapes
    const long long pts = 1LL << lg;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    fprintf(stderr, "device %s, %d SMs, batch 2^%d points\n", prop.name, sms, lg);
    float2 *in, *out, *tw;
    CK(cudaMalloc(&in, pts * 8));
    CK(cudaMalloc(&out, pts * 8 + (2 << 20)));
    fill_kernel<<<sms * 8, 256>>>((float*)in, (size_t)pts * 2, 20260101u);
    std::vector<float2> h(kTwiddleTableSize);
    for (int j = 0; j < kTwiddleTableSize; j++) {
        const double a = -2.0 * M_PI * j / kTwiddleTableSize;
        h[j] = make_float2((float)cos(a), (float)sin(a));
    }
    CK(cudaMalloc(&tw, sizeof(float2) * kTwiddleTableSize));
    CK(cudaMemcpy(tw, h.data(), sizeof(float2) * kTwiddleTableSize, cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    printf("kind,e,n,b,tile_e,stages,minb,io,tw,reorder,reps,hint,out_off,promo,swz,pf,skew,dual,threads,smem,ctas_per_sm,regs,ms_med,ms_min,gbps_med,frac_of_copy,check_rel_l2\n");
    // roofline reference: device copy of the same batch
    double copy_ms = 1e9;
    {
        std::vector<float> t;
        for (int r = 0; r < reps + 2; r++) {
            CK(cudaEventRecord(e0));
            copy_kernel<<<sms * 16, 512>>>((const float4*)in, (float4*)out, (size_t)pts / 2);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        copy_ms = t[t.size() / 2];
        printf("copy_kernel,0,0,0,0,0,0,ldg128,,,,,,,,,,,512,0,16,0,%.4f,%.4f,%.1f,1.000,\n", copy_ms, t[0], pts * 16.0 / copy_ms / 1e6);
        t.clear();
        for (int r = 0; r < reps + 2; r++) {
            CK(cudaEventRecord(e0));
            CK(cudaMemcpyAsync(out, in, pts * 8, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        printf("cudaMemcpyD2D,0,0,0,0,0,0,ce,,,,,,,,,,,0,0,0,0,%.4f,%.4f,%.1f,%.3f,\n", t[t.size() / 2], t[0], pts * 16.0 / t[t.size() / 2] / 1e6, copy_ms / t[t.size() / 2]);
    }
    add_all_sizes(real);
    const size_t CHK = 1 << 18;  // points compared between variants
    std::vector<float2> ref[16][8], got(CHK);
    for (const Variant& v : g_variants) {
        const KernelEntry& k = v.k;
        if (only_e && k.e != only_e) continue;
        cudaFuncAttributes fa;
        if (cudaFuncSetAttribute(k.func, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes) != cudaSuccess) { cudaGetLastError(); continue; }
        {
            // TUNE_CARVEOUT: shared-memory carve-out in percent (100 = max shared, the product setting for the TMA kernels;
            // -1 = driver default).  The LDG/STG paths want the L1 that a small carve-out leaves (tools/fftlike_copy.cu)
            const char* cv = getenv("TUNE_CARVEOUT");
            cudaFuncSetAttribute(k.func, cudaFuncAttributePreferredSharedMemoryCarveout, cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared);
        }
        CK(cudaFuncGetAttributes(&fa, k.func));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.func, k.threads, k.smem_bytes));
        if (per_sm < 1) continue;
        if (v.per_sm > per_sm) continue;  // shape cannot hold that many CTAs per SM
        if (v.per_sm > 0) per_sm = v.per_sm;
        if (v.per_sm < 0) per_sm = 1 << 20;  // one CTA per tile
        TileArgs args;
        memset(&args, 0, sizeof(args));
        // FFT_multiple variants (reps > 1): 1/128 of the batch, transformed reps times in place (CT:669 uses 1/100)
        const long long run_pts = k.reps > 1 ? (pts / 128) / k.tile_points * k.tile_points : pts;
        args.n_points = run_pts;
        args.n_tiles = run_pts / k.tile_points;
        float2* outp = (float2*)((char*)out + v.out_off);
        args.gin = in;
        args.gout = outp;
        args.tw = tw;
        args.l2_hint = v.hint;
        if (io_uses_tma(k.io)) {
            if (encode_tile_map(&args.in_map, in, pts / 16, k.tile_points / 16, v.promo, v.swz) || encode_tile_map(&args.out_map, outp, pts / 16, k.tile_points / 16, v.promo, v.swz)) {
                fprintf(stderr, "tensor map encode failed\n");
                continue;
            }
        }
        long long grid = std::min<long long>((long long)sms * per_sm, args.n_tiles);
        void* params[] = {&args};
        std::vector<float> t;
        CK(cudaMemsetAsync(outp, 0, CHK * 8));
        bool ok = true;
        for (int r = 0; r < reps + 2 && ok; r++) {
            CK(cudaEventRecord(e0));
            cudaError_t le = cudaLaunchKernel(k.func, dim3((unsigned)grid), dim3(k.threads), params, k.smem_bytes, 0);
            CK(cudaEventRecord(e1));
            cudaError_t se = cudaEventSynchronize(e1);
            if (le != cudaSuccess || se != cudaSuccess) {
                fprintf(stderr, "variant failed: %s %s\n", cudaGetErrorString(le), cudaGetErrorString(se));
                ok = false;
                break;
            }
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) t.push_back(ms);
        }
        if (!ok) return 3;  // a failed launch poisons the context: stop
        std::sort(t.begin(), t.end());
        CK(cudaMemcpy(got.data(), outp, CHK * 8, cudaMemcpyDeviceToHost));
        double rel = 0;
        auto& rf = ref[k.e][k.reps > 1 ? 5 + (k.mode != MODE_C2C) : k.mode != MODE_C2C ? 2 + k.mode : (k.reps == 0 ? 2 : k.reorder)];
        if (k.reps == 0 && rf.empty()) {  // staging-only variants must reproduce the input
            rf.resize(CHK);
            CK(cudaMemcpy(rf.data(), in, CHK * 8, cudaMemcpyDeviceToHost));
        }
        if (rf.empty()) {
            rf = got;
        } else {
            double num = 0, den = 0;
            for (size_t i = 0; i < CHK; i++) {
                const double dx = got[i].x - rf[i].x, dy = got[i].y - rf[i].y;
                num += dx * dx + dy * dy;
                den += (double)rf[i].x * rf[i].x + (double)rf[i].y * rf[i].y;
            }
            rel = sqrt(num / den);
        }
        const double med = t[t.size() / 2];
        printf("%s,%d,%d,%d,%d,%d,%d,%s,%s,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%.4f,%.4f,%.1f,%.3f,%.2e\n", k.reps == 0 ? "stage_copy" : (k.mode == MODE_R2C ? "r2c" : k.mode == MODE_C2R ? "c2r" : "fft"), k.e, 1 << k.e, v.b, v.tile_e,
               k.stages, k.minb, k.io == IO_TMA ? "tma" : (k.io == IO_LDG ? "ldg" : k.io == IO_REG ? "reg" : "tma_stg"), k.tw == TW_LUT ? "lut" : "mufu", k.reorder, k.reps, v.hint,
               v.out_off, v.promo, v.swz, k.pf, k.skew, k.dual, k.threads, k.smem_bytes, v.per_sm < 0 ? -1 : per_sm, fa.numRegs, med,
               t[0], pts * 16.0 / med / 1e6, copy_ms / med, rel);
        fflush(stdout);
    }
    return 0;
}
