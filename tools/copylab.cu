// tools/copylab.cu -- what does the B200 memory system want from a streaming read+write kernel?
// (measurement tool, not product).  The staging-only variant of our FFT kernel costs the same as the
// FFT itself (profiles/r01_tune_shapes_c.csv), so the I/O scheme is the ceiling; this program sweeps
// copy schemes over a 4 GiB batch: LDG/STG with different cache operators and unrolls, 1-D TMA bulk
// copies (cp.async.bulk) with different chunk sizes / stages / CTAs per SM / traversal orders / L2
// hints.  Output: CSV rows "scheme,params...,ms_med,ms_min,GBps".
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------------
// LDG/STG copies.  CACHE: 0 default, 1 ld.nc.L1::no_allocate / st.L1::no_allocate, 2 .cs (streaming),
// 3 L1::no_allocate + L2::evict_first policy
template <int CACHE>
__device__ __forceinline__ float4 ld16(const float4* p, uint64_t pol)
{
    float4 r;
    if (CACHE == 0) r = *p;
    if (CACHE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    if (CACHE == 2) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    if (CACHE == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
template <int CACHE>
__device__ __forceinline__ void st16(float4* p, float4 v, uint64_t pol)
{
    if (CACHE == 0) *p = v;
    if (CACHE == 1) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (CACHE == 2) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (CACHE == 3) asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// blocked = 0: grid-stride over 16-byte elements; blocked = 1: each CTA owns one contiguous span
template <int CACHE, int U>
__global__ void __launch_bounds__(1024) ldg_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n, int blocked)
{
    uint64_t pol = 0;
    if (CACHE == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    size_t lo = 0, hi = n, stride = (size_t)gridDim.x * blockDim.x, i;
    if (blocked) {
        const size_t per = (n + gridDim.x - 1) / gridDim.x;
        lo = per * blockIdx.x;
        hi = lo + per < n ? lo + per : n;
        stride = blockDim.x;
        i = lo + threadIdx.x;
    } else {
        i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    }
    for (; i + (U - 1) * stride < hi; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld16<CACHE>(a + i + u * stride, pol);
#pragma unroll
        for (int u = 0; u < U; u++) st16<CACHE>(b + i + u * stride, v[u], pol);
    }
    for (; i < hi; i += stride) st16<CACHE>(b + i, ld16<CACHE>(a + i, pol), pol);
}

// one-shot CTAs (the classic library shape: small CTA, everything loaded up front, huge grid):
// thread t of CTA c copies elements c*blockDim*U + t + u*blockDim; W = 16 or 8 bytes per access
template <int U, int W>
__global__ void oneshot_copy(const char* __restrict__ a, char* __restrict__ b)
{
    extern __shared__ unsigned char pad[];  // only limits CTAs per SM
    const size_t base = ((size_t)blockIdx.x * blockDim.x * U + threadIdx.x) * W;
    if (W == 16) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = *reinterpret_cast<const float4*>(a + base + (size_t)u * blockDim.x * W);
#pragma unroll
        for (int u = 0; u < U; u++) *reinterpret_cast<float4*>(b + base + (size_t)u * blockDim.x * W) = v[u];
    } else {
        float2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = *reinterpret_cast<const float2*>(a + base + (size_t)u * blockDim.x * W);
#pragma unroll
        for (int u = 0; u < U; u++) *reinterpret_cast<float2*>(b + base + (size_t)u * blockDim.x * W) = v[u];
    }
}

// ---------------------------------------------------------------------------------------------
// 1-D TMA bulk copy pipeline: one 32-thread CTA, lane 0 drives everything.
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) bulk_copy(const char* __restrict__ a, char* __restrict__ b, long long n_chunks, int chunk, int stages,
                                                int blocked, int hint, int store_lag)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    if (threadIdx.x != 0) return;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
    uint64_t pol = 0;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int i = 0; i < stages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long first, step, mine;
    if (blocked) {
        const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
        first = per * blockIdx.x;
        step = 1;
        mine = first >= n_chunks ? 0 : (first + per <= n_chunks ? per : n_chunks - first);
    } else {
        first = blockIdx.x;
        step = gridDim.x;
        mine = first < n_chunks ? (n_chunks - first + step - 1) / step : 0;
    }
    auto load = [&](long long k) {
        const int st = (int)(k % stages);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(chunk) : "memory");
        const char* src = a + (first + k * step) * (long long)chunk;
        if (hint & 1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(smem + (size_t)st * chunk)), "l"(src), "r"(chunk), "r"(s32(&full[st])), "l"(pol) : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)st * chunk)), "l"(src), "r"(chunk), "r"(s32(&full[st])) : "memory");
    };
    // loads run `stages - store_lag` chunks ahead; a buffer is refilled once its store has read it
    const int ahead = stages - store_lag;
    for (long long k = 0; k < ahead && k < mine; k++) load(k);
    for (long long k = 0; k < mine; k++) {
        const int st = (int)(k % stages);
        const uint32_t parity = (uint32_t)((k / stages) & 1);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(&full[st])), "r"(parity) : "memory");
        }
        char* dst = b + (first + k * step) * (long long)chunk;
        if (hint & 2)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(s32(smem + (size_t)st * chunk)), "r"(chunk), "l"(pol) : "memory");
        else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(smem + (size_t)st * chunk)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        const long long kn = k + ahead;
        if (kn < mine) {
            // buffer kn % stages was stored (k - store_lag + ... ) steps ago: allow store_lag-1 stores still reading
            // the buffer being refilled belonged to chunk k - store_lag: only the store_lag newest stores may still be reading
            if (store_lag == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            load(kn);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// same single-thread pipeline, but with 2-D tensor-map TMA (what the FFT kernels use):
// the batch is rows of `inner` bytes; a chunk is a box of chunk/inner rows.
__global__ void __launch_bounds__(32) tensor_copy(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                                                  long long n_chunks, int chunk, int rows_per_chunk, int stages)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    if (threadIdx.x != 0) return;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
    for (int i = 0; i < stages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long first = blockIdx.x, step = gridDim.x;
    const long long mine = first < n_chunks ? (n_chunks - first + step - 1) / step : 0;
    auto load = [&](long long k) {
        const int st = (int)(k % stages);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(smem + (size_t)st * chunk)),
                     "l"(reinterpret_cast<uint64_t>(&in_map)), "r"(0), "r"((int)((first + k * step) * rows_per_chunk)), "r"(s32(&full[st])) : "memory");
    };
    const int ahead = stages - 1;
    for (long long k = 0; k < ahead && k < mine; k++) load(k);
    for (long long k = 0; k < mine; k++) {
        const int st = (int)(k % stages);
        const uint32_t parity = (uint32_t)((k / stages) & 1);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(&full[st])), "r"(parity) : "memory");
        }
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(&out_map)), "r"(0),
                     "r"((int)((first + k * step) * rows_per_chunk)), "r"(s32(smem + (size_t)st * chunk)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        const long long kn = k + ahead;
        if (kn < mine) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(kn);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(void* base, size_t bytes, int inner_bytes, int box_rows, bool swz)
{
    static EncFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        enc = (EncFn)p;
    }
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)inner_bytes / 4, (cuuint64_t)(bytes / inner_bytes)};
    cuuint64_t gstr[1] = {(cuuint64_t)inner_bytes};
    cuuint32_t box[2] = {(cuuint32_t)inner_bytes / 4, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d (inner %d box_rows %d)\n", (int)r, inner_bytes, box_rows); exit(4); }
    return m;
}

static float time_med(std::vector<float>& t, float* mn)
{
    std::sort(t.begin(), t.end());
    *mn = t[0];
    return t[t.size() / 2];
}

int main(int argc, char** argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 32;  // log2 bytes
    const int reps = argc > 2 ? atoi(argv[2]) : 7;
    const size_t bytes = (size_t)1 << lg;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    char *a, *b;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes));
    CK(cudaMemset(b, 0, bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("scheme,p1,p2,p3,p4,p5,p6,ms_med,ms_min,GBps\n");
    auto run = [&](const char* name, int p1, int p2, int p3, int p4, int p5, int p6, auto&& launch) {
        std::vector<float> t;
        for (int r = 0; r < reps + 2; r++) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { fprintf(stderr, "%s failed: %s\n", name, cudaGetErrorString(e)); exit(3); }
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) t.push_back(ms);
        }
        float mn, med = time_med(t, &mn);
        printf("%s,%d,%d,%d,%d,%d,%d,%.4f,%.4f,%.1f\n", name, p1, p2, p3, p4, p5, p6, med, mn, 2.0 * bytes / med / 1e6);
        fflush(stdout);
    };
    run("cudaMemcpyD2D", 0, 0, 0, 0, 0, 0, [&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); });
    const size_t n16 = bytes / 16;
    // LDG/STG: cache op x unroll x block x CTAs/SM x traversal
#define LDGV(C, U)                                                                                             \
    for (int blk : {256, 512, 1024})                                                                            \
        for (int per : {1, 2, 4, 8})                                                                            \
            for (int blocked : {0, 1}) {                                                                        \
                if (blk * per > 2048) continue;                                                                 \
                run("ldg_copy", C, U, blk, per, blocked, 0,                                                     \
                    [&] { ldg_copy<C, U><<<sms * per, blk>>>((const float4*)a, (float4*)b, n16, blocked); });   \
            }
    if (argc > 3) { LDGV(1, 2) LDGV(1, 4) }
    // one-shot CTAs: U x width x block x CTAs/SM (limited through dynamic smem)
#define ONES(U, W)                                                                                                 \
    for (int blk : {64, 128, 256})                                                                                 \
        for (int per : {2, 4, 8, 12, 16, 24, 32}) {                                                                \
            if (blk * per > 2048) continue;                                                                        \
            const size_t smem = (220 * 1024 / per) & ~1023;                                                        \
            CK(cudaFuncSetAttribute(oneshot_copy<U, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            run("oneshot_copy", U, W, blk, per, 0, 0,                                                              \
                [&] { oneshot_copy<U, W><<<(unsigned)(bytes / ((size_t)blk * U * W)), blk, smem>>>(a, b); });      \
        }
    if (argc > 3) { ONES(4, 16) ONES(8, 16) ONES(16, 16) ONES(8, 8) ONES(16, 8) ONES(32, 8) }
    // 2-D tensor TMA: inner row bytes x swizzle x chunk x stages x CTAs/SM
    CK(cudaFuncSetAttribute(tensor_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int inner : {128, 256, 512, 1024})
        for (int swz : {1, 0}) {
            if (swz && inner != 128) continue;
            for (int chunk : {8192, 16384, 32768})
                for (int stages : {2, 3, 4, 5, 6})
                    for (int per : {1, 2, 3, 4, 6}) {
                        if (inner != 128 || !swz) continue;
                        const size_t smem = (size_t)chunk * stages + 8 * stages + 1024;
                        if (smem * per > 220 * 1024 || chunk / inner > 256) continue;
                        CUtensorMap im = make_map(a, bytes, inner, chunk / inner, swz), om = make_map(b, bytes, inner, chunk / inner, swz);
                        run("tensor_copy", inner, swz, chunk, stages, per, 0, [&] {
                            tensor_copy<<<sms * per, 32, smem>>>(im, om, (long long)(bytes / chunk), chunk, chunk / inner, stages);
                        });
                    }
        }
    // TMA bulk: chunk x stages x CTAs/SM x traversal x hint x store_lag
    CK(cudaFuncSetAttribute(bulk_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int chunk : {32768})
        for (int stages : {3})
            for (int per : {1, 2})
                for (int blocked : {0, 1})
                    for (int hint : {0, 3})
                        for (int lag : {1, 2}) {
                            const size_t smem = (size_t)chunk * stages + 8 * stages;
                            if (smem * per > 220 * 1024 || lag >= stages) continue;
                            if ((size_t)chunk * stages * per < 48 * 1024) continue;  // too little in flight to matter
                            run("bulk_copy", chunk, stages, per, blocked, hint, lag, [&] {
                                bulk_copy<<<sms * per, 32, smem>>>(a, b, (long long)(bytes / chunk), chunk, stages, blocked, hint, lag);
                            });
                        }
    return 0;
}
