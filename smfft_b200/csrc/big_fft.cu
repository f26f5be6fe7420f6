// big_fft.cu -- C2C transforms of 2^15 .. 2^18 points: two passes over HBM (the "four-step" factorisation).
//
// Beyond the reference ("N > 4096 via multi-pass"): KAdamek/SMFFT stops where one transform stops fitting one
// CTA's shared memory.  N = N1 * N2, n = n1 + N1 n2, k = N2 k1 + k2:
//
//   X[N2 k1 + k2] = sum_n1 W_N1^(n1 k1) * [ W_N^(n1 k2) * sum_n2 x[n1 + N1 n2] W_N2^(n2 k2) ]
//
//   pass A: for every n1 a transform of length N2 over n2 (stride N1 in memory), times W_N^(n1 k2), written back to the same
//           positions (n1 + N1 k2) of a scratch buffer.  A CTA owns 16 consecutive n1: its tile is a TMA box of N2 rows x 128
//           bytes with a row pitch of N1 * 8 bytes -- strided in HBM, dense (SWIZZLE_128B) in shared memory.
//   pass B: for every k2 a transform of length N1 over n1 (contiguous), written to X[N2 k1 + k2].  A CTA owns 16 consecutive
//           k2: it reads 16 contiguous transforms (one TMA box of N1 rows x 128 bytes, dense) and writes a TMA box of N1 rows
//           x 128 bytes with a row pitch of N2 * 8 bytes.  (Reading the rows straight into registers instead: 2.81 / 2.87 ms
//           against 2.65 / 2.62 ms at 2^15 / 2^16 points, profiles/r02_ab_two_pass.json.)
//
// Both passes are user kernels of the library's own device primitive (smfft::BlockFFT, include/smfft/device.cuh): 16
// transforms per block, 16 points per thread, registers in and out; the shared-memory tile is the TMA landing zone, then the
// primitive's exchange scratch, then the transposed staging of the TMA store.  Column c of row r of a box sits at
// LayoutSW128::phys(16 r + c): the 16 lanes of a transform read 16 rows of one column and, with two adjacent transforms per
// warp, touch every bank once.
// Algorithmic traffic: 32 bytes per point (each pass reads and writes the batch once); the batch is processed in chunks so
// that the scratch stays small (and, for small chunks, resident in the 126 MB L2 between the passes).
#include "big_fft.hpp"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "smfft/detail/tma.cuh"
#include "smfft/device.cuh"
#include "tmap.hpp"

namespace smfft {
namespace big {

namespace {

#ifndef SMFFT_BIG_R32_512
#define SMFFT_BIG_R32_512 1  // 512-point passes: 32 points per thread (16 lanes per transform) instead of 16 (32 lanes): 2.95 / 3.08 vs 3.16 / 3.58 ms;
                             // persistent CTAs with three 64 KB buffers instead of one tile per CTA: 2.96 / 3.11 ms, no gain (profiles/r02_ab_two_pass.json)
#endif
#ifndef SMFFT_BIG_SPLIT17
#define SMFFT_BIG_SPLIT17 8  // log2 of the strided pass (A) at 2^17 points: 8 (256 x 512) or 9 (512 x 256)
#endif
SMFFT_CX int log2r_for(int log2len) { return (log2len >= 9 && SMFFT_BIG_R32_512) ? 5 : 4; }

struct PassArgs {
    alignas(64) CUtensorMap in_map;   // pass A: [N2 * ffts rows][N1 points]; pass B: the scratch as rows of 128 bytes
    alignas(64) CUtensorMap out_map;  // pass A: the same geometry over the scratch; pass B: [N1 * ffts rows][N2 points]
    const float2* base_tw;            // W_16384 table (twiddles of the block transforms)
    const float2* wt;                 // pass A: W_N^j, j < 512, then W_N^(512 j), j < N / 512
    int groups;                       // tiles per transform: N1 / 16 (pass A), N2 / 16 (pass B)
};

// one tile = 16 transforms of 2^LOG2LEN points.  PASS 0 = A (strided box in, twiddle, same box out), 1 = B (16 contiguous
// transforms in -- in_map: the scratch as rows of 128 bytes --, strided box out)
template <int LOG2LEN, int DIR, int PASS>
__global__ void __launch_bounds__((16 << (LOG2LEN - log2r_for(LOG2LEN))), (LOG2LEN <= 7 ? 8 : LOG2LEN == 8 ? 4 : 2)) big_pass_kernel(const __grid_constant__ PassArgs a)
{
    using F = BlockFFT<LOG2LEN, DIR, 16, TW_LUT, log2r_for(LOG2LEN)>;  // 512 points: 32 per thread, so a transform keeps 16 lanes
    using SW = detail::LayoutSW128;
    constexpr int LEN = 1 << LOG2LEN, TILE = 16 * LEN, T = F::T, R = F::R;
    constexpr int BOX_ROWS = LEN > 256 ? 256 : LEN, NBOX = LEN / BOX_ROWS;
    static_assert(F::THREADS == 16 * T, "16 transforms per block");
    extern __shared__ unsigned char raw[];
    unsigned char* smem = raw + ((1024u - (plat::smem_u32(raw) & 1023u)) & 1023u);  // SWIZZLE_128B needs a 1 KB aligned tile
    float2* tile = reinterpret_cast<float2*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TILE * 8);
    float2* stw = reinterpret_cast<float2*>(smem + TILE * 8 + 64);
    const int tid = threadIdx.x, f = tid / T, t = tid & (T - 1);
    const long long id = blockIdx.x;
    const int fft = (int)(id / a.groups), g = (int)(id % a.groups);
    const int row0 = fft * LEN;  // first row of this transform in the strided matrix

    float2 v[R];
    if (tid == 0) {
        plat::mbar_init(bar, 1);
        plat::mbar_fence_init();
        plat::mbar_arrive_expect_tx(bar, TILE * 8);
#pragma unroll
        for (int b = 0; b < NBOX; b++) {
            if constexpr (PASS == 0)
                plat::tma_load_2d(tile + b * BOX_ROWS * 16, &a.in_map, 32 * g, row0 + b * BOX_ROWS, bar);
            else
                plat::tma_load_2d(tile + b * BOX_ROWS * 16, &a.in_map, 0, (int)(id * LEN) + b * BOX_ROWS, bar);  // 16 contiguous transforms = LEN rows of 128 bytes
        }
    }
    F::fill_twiddles(stw, a.base_tw);
    __syncthreads();  // barrier initialised, table filled
    plat::mbar_wait(bar, 0);
    if constexpr (PASS == 0) {
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = tile[SW::phys((t + m * T) * 16 + f)];  // column f of the box
    } else {
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = tile[SW::phys(F::index(m))];  // transform f of the tile: 16 lanes read one 128-byte row
    }

    F::exec(v, tile, stw);  // synchronises before its first write to the tile: every thread has its points in registers

    if constexpr (PASS == 0) {
        // W_N^(n1 k2), k2 = t + m T = W^(n1 t) (W^(n1 T))^m: an accurate base and an accurate step from the two-level table,
        // the powers four at a time (at most R/4 + 2 roundings deep)
        const int n1 = 16 * g + f;
        auto W = [&](int p) {
            float2 w = detail::cmul(__ldg(a.wt + (p & 511)), __ldg(a.wt + 512 + (p >> 9)));
            if (DIR) w.y = -w.y;
            return w;
        };
        const float2 s1 = W(n1 * T), s2 = detail::csqr(s1), s3 = detail::cmul(s2, s1), s4 = detail::csqr(s2);
        float2 bq = W(n1 * t);
#pragma unroll
        for (int q = 0; q < R / 4; q++) {
            v[4 * q] = detail::cmul(v[4 * q], bq);
            v[4 * q + 1] = detail::cmul(v[4 * q + 1], detail::cmul(bq, s1));
            v[4 * q + 2] = detail::cmul(v[4 * q + 2], detail::cmul(bq, s2));
            v[4 * q + 3] = detail::cmul(v[4 * q + 3], detail::cmul(bq, s3));
            if (q + 1 < R / 4) bq = detail::cmul(bq, s4);
        }
    }

    __syncthreads();  // the last exchange has been read by every thread
#pragma unroll
    for (int m = 0; m < R; m++) tile[SW::phys((t + m * T) * 16 + f)] = v[m];
    plat::fence_proxy_async();  // generic-proxy writes -> TMA store (async proxy)
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBOX; b++) plat::tma_store_2d(&a.out_map, 32 * g, row0 + b * BOX_ROWS, tile + b * BOX_ROWS * 16);
        plat::bulk_commit();
        plat::bulk_wait_read0();  // the tile must outlive the store's reads
    }
}

template <int LOG2LEN>
constexpr int pass_smem_bytes()
{
    return 16 * (1 << LOG2LEN) * 8 + 64 + ((BlockFFT<LOG2LEN, 0, 16, TW_LUT, log2r_for(LOG2LEN)>::TWIDDLE_POINTS * 8 + 127) & ~127) + 1024;
}

typedef void (*PassFn)(const PassArgs);

struct PassInfo {
    PassFn fn;
    int threads, smem;
};

template <int LOG2LEN>
PassInfo pass_info(int dir, int pass)
{
    PassFn fn = dir ? (pass ? big_pass_kernel<LOG2LEN, 1, 1> : big_pass_kernel<LOG2LEN, 1, 0>)
                    : (pass ? big_pass_kernel<LOG2LEN, 0, 1> : big_pass_kernel<LOG2LEN, 0, 0>);
    return PassInfo{fn, 16 << (LOG2LEN - log2r_for(LOG2LEN)), pass_smem_bytes<LOG2LEN>()};
}

PassInfo pass_for(int log2len, int dir, int pass)
{
    switch (log2len) {
        case 7: return pass_info<7>(dir, pass);
        case 8: return pass_info<8>(dir, pass);
        default: return pass_info<9>(dir, pass);
    }
}

struct DevState {
    std::mutex mu;
    float2* wt[kMaxLog2 + 1] = {};
    cudaMemPool_t pool = nullptr;
    std::vector<const void*> attr_done;
};
DevState g_state[64];

int failf(char* err, int cap, int* cuda, int code, const char* fmt, const char* what)
{
    if (err && cap > 0) snprintf(err, cap, fmt, what);
    if (cuda) *cuda = code;
    return 1;
}

#define BIG_TRY(expr)                                                                                            \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            return failf(err, errcap, cuda, e__ == cudaErrorMemoryAllocation ? 2 : 1, "smfft (two-pass transform): " #expr ": %s", cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

// the factorisation: N2 = length of pass A (strided), N1 = length of pass B (contiguous)
static void split(int e, int* log2_n2, int* log2_n1)
{
    *log2_n2 = e == 18 ? 9 : e == 17 ? SMFFT_BIG_SPLIT17 : 8;
    *log2_n1 = e - *log2_n2;
}

int exec(const Params& p, long long* launches, char* err, int errcap, int* cuda)
{
    if (p.e < kMinLog2 || p.e > kMaxLog2) return failf(err, errcap, cuda, 0, "smfft: two-pass transforms cover 2^15 .. 2^18 points%s", "");
    if (p.n_ffts <= 0) return 0;
    int dev = -1;
    BIG_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return failf(err, errcap, cuda, 0, "smfft: device ordinal out of range%s", "");
    DevState& st = g_state[dev];
    int l2, l1;
    split(p.e, &l2, &l1);
    const long long N = 1LL << p.e, N1 = 1LL << l1, N2 = 1LL << l2;
    const PassInfo pa = pass_for(l2, p.dir, 0), pb = pass_for(l1, p.dir, 1);
    {
        std::lock_guard<std::mutex> lk(st.mu);
        if (!st.wt[p.e]) {
            // W_N^j for j < 512 and W_N^(512 j) for j < N / 512 (<= 512), forward sign, rounded from FP64
            std::vector<float2> h(1024, make_float2(1.0f, 0.0f));
            for (int j = 0; j < 512; j++) {
                const double a0 = -2.0 * M_PI * (double)j / (double)N, a1 = -2.0 * M_PI * (double)j * 512.0 / (double)N;
                h[j] = make_float2((float)cos(a0), (float)sin(a0));
                if (j < N / 512) h[512 + j] = make_float2((float)cos(a1), (float)sin(a1));
            }
            float2* d = nullptr;
            BIG_TRY(cudaMalloc((void**)&d, sizeof(float2) * 1024));
            cudaError_t e = cudaMemcpy(d, h.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                cudaFree(d);
                return failf(err, errcap, cuda, 1, "smfft: twiddle table upload failed: %s", cudaGetErrorString(e));
            }
            st.wt[p.e] = d;
        }
        if (!st.pool) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            BIG_TRY(cudaMemPoolCreate(&st.pool, &props));
            unsigned long long keep = ~0ULL;  // the scratch of one call serves the next: nothing goes back to the driver until release()
            BIG_TRY(cudaMemPoolSetAttribute(st.pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        for (const PassInfo* k : {&pa, &pb}) {
            bool need = true;
            for (const void* f : st.attr_done) need &= (f != (const void*)k->fn);
            if (need) {
                BIG_TRY(cudaFuncSetAttribute((const void*)k->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k->smem));
                BIG_TRY(cudaFuncSetAttribute((const void*)k->fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
                st.attr_done.push_back((const void*)k->fn);
            }
        }
    }
    const long long fft_bytes = N * (long long)sizeof(float2);
    long long chunk = p.chunk_bytes / fft_bytes;
    if (chunk < 1) chunk = 1;
    if (chunk > p.n_ffts) chunk = p.n_ffts;
    float2* scratch = nullptr;
    BIG_TRY(cudaMallocFromPoolAsync((void**)&scratch, (size_t)(chunk * fft_bytes), st.pool, p.stream));  // stream-ordered: safe across host threads and streams
    int rc = 0;
    for (long long f0 = 0; f0 < p.n_ffts && !rc; f0 += chunk) {
        const long long cf = p.n_ffts - f0 < chunk ? p.n_ffts - f0 : chunk;
        const float2* in = (const float2*)p.in + f0 * N;
        float2* out = (float2*)p.out + f0 * N;
        PassArgs a;
        memset(&a, 0, sizeof(a));
        a.base_tw = (const float2*)p.base_tw;
        a.wt = st.wt[p.e];
        // pass A: x -> scratch
        int r0 = host::encode_strided_map(&a.in_map, in, 2 * N1, N2 * cf, N2 > 256 ? 256 : (int)N2);
        int r1 = host::encode_strided_map(&a.out_map, scratch, 2 * N1, N2 * cf, N2 > 256 ? 256 : (int)N2);
        if (r0 || r1) { rc = failf(err, errcap, cuda, 1, "smfft: cuTensorMapEncodeTiled failed (two-pass transform, pass A)%s", ""); break; }
        a.groups = (int)(N1 / 16);
        void* params[] = {&a};
        cudaError_t e = cudaLaunchKernel((const void*)pa.fn, dim3((unsigned)(cf * a.groups)), dim3((unsigned)pa.threads), params, (size_t)pa.smem, p.stream);
        if (e != cudaSuccess) { rc = failf(err, errcap, cuda, 1, "smfft: two-pass transform, pass A launch: %s", cudaGetErrorString(e)); break; }
        // pass B: scratch -> X
        PassArgs b;
        memset(&b, 0, sizeof(b));
        b.base_tw = (const float2*)p.base_tw;
        if (host::encode_tile_map(&b.in_map, scratch, cf * N / 16, N1 > 256 ? 256 : (int)N1)) {
            rc = failf(err, errcap, cuda, 1, "smfft: cuTensorMapEncodeTiled failed (two-pass transform, pass B input)%s", "");
            break;
        }
        r1 = host::encode_strided_map(&b.out_map, out, 2 * N2, N1 * cf, N1 > 256 ? 256 : (int)N1);
        if (r1) { rc = failf(err, errcap, cuda, 1, "smfft: cuTensorMapEncodeTiled failed (two-pass transform, pass B)%s", ""); break; }
        b.groups = (int)(N2 / 16);
        void* params_b[] = {&b};
        e = cudaLaunchKernel((const void*)pb.fn, dim3((unsigned)(cf * b.groups)), dim3((unsigned)pb.threads), params_b, (size_t)pb.smem, p.stream);
        if (e != cudaSuccess) { rc = failf(err, errcap, cuda, 1, "smfft: two-pass transform, pass B launch: %s", cudaGetErrorString(e)); break; }
        if (launches) *launches += 2;
    }
    cudaError_t e = cudaFreeAsync(scratch, p.stream);
    if (!rc && e != cudaSuccess) rc = failf(err, errcap, cuda, 1, "smfft: cudaFreeAsync: %s", cudaGetErrorString(e));
    return rc;
}

void release()
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    DevState& st = g_state[dev];
    std::lock_guard<std::mutex> lk(st.mu);
    for (float2*& w : st.wt) {
        if (w) cudaFree(w);
        w = nullptr;
    }
    if (st.pool) {
        cudaDeviceSynchronize();
        cudaMemPoolDestroy(st.pool);
        st.pool = nullptr;
    }
}

}  // namespace big
}  // namespace smfft
