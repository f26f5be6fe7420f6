// big_fft.cu -- C2C transforms of 2^15 .. 2^24 points: two passes over HBM (the "four-step" factorisation), three from 2^21.
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
    alignas(64) CUtensorMap in_map;   // column pass: [rows][row_len points]; row pass: the scratch as rows of 128 bytes
    alignas(64) CUtensorMap out_map;  // column pass: the same geometry over the destination; row pass: [LEN * ffts rows][out_row_len points]
    const float2* base_tw;            // W_16384 table (twiddles of the block transforms)
    const float2* wt;                 // column pass: W_M^j (j < 512), W_M^(512 j), W_M^(2^18 j) -- M = the modulus of this pass's twiddles
    int groups;                       // column pass: 16-column groups per row; row pass: 16-transform groups per (transform, kmid)
    // column pass: tile id -> slab = id / groups (LEN consecutive rows), g = id % groups; the value at (row k, column c) is
    // multiplied by W_M^(cidx (kfix + kscale k)), cidx = c >> col_shift, kfix = slab % kdiv
    int col_shift, kdiv, kscale;
    // row pass: tile id -> g = id % groups, kmid = (id / groups) % nmid, fft = id / (groups nmid); transform j of the tile is
    // transform kmid + nmid (16 g + j + 16 groups fft) of the scratch; column offset of the output box = out_col_scale * kmid
    int nmid, out_col_scale;
};

// one tile = 16 transforms of 2^LOG2LEN points.  PASS 0 = column pass (strided box in, twiddle, same box out), 1 = row pass
// (16 transforms in -- contiguous, or gathered one box each when nmid > 1 --, strided box out)
template <int LOG2LEN, int DIR, int PASS, int LOG2W>
__global__ void __launch_bounds__(((1 << LOG2W) << (LOG2LEN - log2r_for(LOG2LEN))), (LOG2LEN <= 7 ? 8 : LOG2LEN == 8 ? 4 : 2)) big_pass_kernel(const __grid_constant__ PassArgs a)
{
    // WD transforms per tile: 16 (one 128-byte line per box row), or 8 for 1024-point passes (64-byte segments, 64 KB tiles)
    constexpr int WD = 1 << LOG2W;
    using F = BlockFFT<LOG2LEN, DIR, WD, TW_LUT, log2r_for(LOG2LEN)>;  // 512 points and up: 32 per thread
    using SW = detail::LayoutSW128;  // tiles that arrive as rows of 128 bytes (the row pass's input)
    // strided boxes: rows of WD points.  16 points = 128 bytes: SWIZZLE_128B.  8 points = 64 bytes: SWIZZLE_64B (16-byte chunk
    // index, bits 4-5 of the address, XOR bits 7-8), two box rows per 128-byte line
    struct SW64 {
        static __device__ __forceinline__ int phys(int x) { return x ^ (((x >> 4) & 3) << 1); }
    };
    using SWT = typename std::conditional<WD == 8, SW64, SW>::type;
    constexpr int LEN = 1 << LOG2LEN, TILE = WD * LEN, T = F::T, R = F::R;
    constexpr int BOX_ROWS = LEN > 256 ? 256 : LEN, NBOX = LEN / BOX_ROWS;            // strided boxes: LEN rows of WD points
    constexpr int IN_ROWS = TILE / 16, IN_BOX = IN_ROWS > 256 ? 256 : IN_ROWS, IN_NBOX = IN_ROWS / IN_BOX;  // contiguous tile: rows of 128 bytes
    static_assert(F::THREADS == WD * T, "WD transforms per block");
    extern __shared__ unsigned char raw[];
    unsigned char* smem = raw + ((1024u - (plat::smem_u32(raw) & 1023u)) & 1023u);  // SWIZZLE_128B needs a 1 KB aligned tile
    float2* tile = reinterpret_cast<float2*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TILE * 8);
    float2* stw = reinterpret_cast<float2*>(smem + TILE * 8 + 64);
    const int tid = threadIdx.x, f = tid / T, t = tid & (T - 1);
    const unsigned id = blockIdx.x;  // tile ids fit 32 bits (the grid does): 32-bit divisions only
    const int g = (int)(id % (unsigned)a.groups);
    const unsigned slab = id / (unsigned)a.groups;  // column pass: which LEN rows; row pass: fft * nmid + kmid
    int out_c0, out_row0;
    if constexpr (PASS == 0) {
        out_c0 = 2 * WD * g;
        out_row0 = (int)(slab * LEN);
    } else {
        const int kmid = (int)(slab % (unsigned)a.nmid);
        out_c0 = 2 * WD * g + 2 * a.out_col_scale * kmid;
        out_row0 = (int)(slab / (unsigned)a.nmid) * LEN;
    }

    float2 v[R];
    if (tid == 0) {
        plat::mbar_init(bar, 1);
        plat::mbar_fence_init();
    }
    __syncthreads();  // the barrier is initialised for every thread (and for the tools) before its first use
    if (tid == 0) {
        plat::mbar_arrive_expect_tx(bar, TILE * 8);
        if constexpr (PASS == 0) {
#pragma unroll
            for (int b = 0; b < NBOX; b++) plat::tma_load_2d(tile + b * BOX_ROWS * WD, &a.in_map, out_c0, out_row0 + b * BOX_ROWS, bar);
        } else if (a.nmid == 1) {
#pragma unroll
            for (int b = 0; b < IN_NBOX; b++) plat::tma_load_2d(tile + b * IN_BOX * 16, &a.in_map, 0, (int)(id * (unsigned)IN_ROWS) + b * IN_BOX, bar);  // WD contiguous transforms = IN_ROWS rows of 128 bytes
        } else {
            // gathered: transform j of the tile is LEN / 16 rows of 128 bytes somewhere in the scratch, one box each
            const unsigned fft = slab / (unsigned)a.nmid, kmid = slab % (unsigned)a.nmid;
            const long long tr0 = kmid + (long long)a.nmid * (WD * g + (long long)WD * a.groups * fft);
            for (int j = 0; j < WD; j++) plat::tma_load_2d(tile + j * LEN, &a.in_map, 0, (int)((tr0 + (long long)a.nmid * j) * (LEN / 16)), bar);
        }
    }
    F::fill_twiddles(stw, a.base_tw);
    __syncthreads();  // table filled
    plat::mbar_wait(bar, 0);
    if constexpr (PASS == 0) {
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = tile[SWT::phys((t + m * T) * WD + f)];  // column f of the box
    } else {
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = tile[SW::phys(F::index(m))];  // transform f of the tile: 16 lanes read one 128-byte row
    }

    F::exec(v, tile, stw);  // synchronises before its first write to the tile: every thread has its points in registers

    if constexpr (PASS == 0) {
        // W_M^(cidx (kfix + kscale k)), k = t + m T: an accurate base and an accurate step from the three-level table, the
        // powers four at a time (at most R/4 + 2 roundings deep)
        const unsigned cidx = (unsigned)(WD * g + f) >> a.col_shift;
        auto W = [&](unsigned p) {  // p < M <= 2^24
            float2 w = detail::cmul(__ldg(a.wt + (p & 511)), __ldg(a.wt + 512 + ((p >> 9) & 511)));
            w = detail::cmul(w, __ldg(a.wt + 1024 + (p >> 18)));
            if (DIR) w.y = -w.y;
            return w;
        };
        const unsigned kfix = slab % (unsigned)a.kdiv;
        const float2 s1 = W(cidx * (unsigned)(a.kscale * T)), s2 = detail::csqr(s1), s3 = detail::cmul(s2, s1), s4 = detail::csqr(s2);
        float2 bq = W(cidx * (kfix + (unsigned)a.kscale * t));
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
    for (int m = 0; m < R; m++) tile[SWT::phys((t + m * T) * WD + f)] = v[m];
    plat::fence_proxy_async();  // generic-proxy writes -> TMA store (async proxy)
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBOX; b++) plat::tma_store_2d(&a.out_map, out_c0, out_row0 + b * BOX_ROWS, tile + b * BOX_ROWS * WD);
        plat::bulk_commit();
        plat::bulk_wait_read0();  // the tile must outlive the store's reads
    }
}

#ifndef SMFFT_BIG_W8_FROM
#define SMFFT_BIG_W8_FROM 9  // 512-point passes too: 32 KB tiles, four CTAs per SM instead of two (2^17: 2.75 -> 2.64 ms, 2^18: 2.97 -> 2.78 ms)
#endif
SMFFT_CX int log2w_for(int log2len) { return log2len >= SMFFT_BIG_W8_FROM ? 3 : 4; }  // transforms per tile: 16, or 8 for 512- and 1024-point passes

template <int LOG2LEN>
constexpr int pass_smem_bytes()
{
    return (1 << log2w_for(LOG2LEN)) * (1 << LOG2LEN) * 8 + 64 +
           ((BlockFFT<LOG2LEN, 0, (1 << log2w_for(LOG2LEN)), TW_LUT, log2r_for(LOG2LEN)>::TWIDDLE_POINTS * 8 + 127) & ~127) + 1024;
}

typedef void (*PassFn)(const PassArgs);

struct PassInfo {
    PassFn fn;
    int threads, smem;
    int width;  // transforms per tile
};

template <int LOG2LEN>
PassInfo pass_info(int dir, int pass)
{
    constexpr int LW = log2w_for(LOG2LEN);
    PassFn fn = dir ? (pass ? big_pass_kernel<LOG2LEN, 1, 1, LW> : big_pass_kernel<LOG2LEN, 1, 0, LW>)
                    : (pass ? big_pass_kernel<LOG2LEN, 0, 1, LW> : big_pass_kernel<LOG2LEN, 0, 0, LW>);
    return PassInfo{fn, (1 << LW) << (LOG2LEN - log2r_for(LOG2LEN)), pass_smem_bytes<LOG2LEN>(), 1 << LW};
}

PassInfo pass_for(int log2len, int dir, int pass)
{
    switch (log2len) {
        case 6: return pass_info<6>(dir, pass);
        case 7: return pass_info<7>(dir, pass);
        case 8: return pass_info<8>(dir, pass);
        case 9: return pass_info<9>(dir, pass);
        default: return pass_info<10>(dir, pass);
    }
}

struct DevState {
    std::mutex mu;
    float2* wt[kMaxLog2 + 1] = {};  // three-level twiddle tables by log2 of the modulus
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
            return failf(err, errcap, cuda, e__ == cudaErrorMemoryAllocation ? 2 : 1, "smfft (multi-pass transform): " #expr ": %s", cudaGetErrorString(e__)); \
    } while (0)

// W_M^j (j < 512), W_M^(512 j), W_M^(2^18 j), forward sign, rounded from FP64: W_M^p = lo[p & 511] mid[(p >> 9) & 511] hi[p >> 18]
int twiddle_table(DevState& st, int log2m, float2** out, char* err, int errcap, int* cuda)
{
    if (!st.wt[log2m]) {
        const long long M = 1LL << log2m;
        std::vector<float2> h(1536);
        for (int lvl = 0; lvl < 3; lvl++)
            for (long long j = 0; j < 512; j++) {
                const long long p = (j << (lvl == 0 ? 0 : lvl == 1 ? 9 : 18)) % M;
                const double ang = -2.0 * M_PI * (double)p / (double)M;
                h[lvl * 512 + j] = make_float2((float)cos(ang), (float)sin(ang));
            }
        float2* d = nullptr;
        BIG_TRY(cudaMalloc((void**)&d, sizeof(float2) * h.size()));
        cudaError_t e = cudaMemcpy(d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            return failf(err, errcap, cuda, 1, "smfft: twiddle table upload failed: %s", cudaGetErrorString(e));
        }
        st.wt[log2m] = d;
    }
    *out = st.wt[log2m];
    return 0;
}

}  // namespace

// The factorisation.  Two passes (2^15 .. 2^20): N = N1 N2, lengths (N2 strided, N1 contiguous).  Three passes (2^21 .. 2^24):
// N = N1 N2 N3, n = n1 + N1 n2 + N1 N2 n3, k = k3 + N3 k2 + N2 N3 k1:
//   pass 1: over n3 (stride N1 N2), times W_(N2 N3)^(n2 k3);  pass 2: over n2 (stride N1), times W_N^(n1 (k3 + N3 k2));
//   pass 3: over n1 (contiguous), written to X[k3 + N3 k2 + N2 N3 k1] -- a tile gathers 16 consecutive k3.
// All three-pass lengths are 64 .. 256 points (32 KB tiles at most).
static void split(int e, int* l3, int* l2, int* l1)
{
    if (e <= 20) {
        // two passes; 2^19 = 512 x 1024 and 2^20 = 1024 x 1024 run 1024-point passes on 8-transform tiles (64 KB)
        *l3 = 0;
        *l2 = e == 20 ? 10 : e >= 18 ? 9 : e == 17 ? SMFFT_BIG_SPLIT17 : 8;
        *l1 = e - *l2;
    } else {
        *l1 = (e + 2) / 3;
        *l2 = (e - *l1 + 1) / 2;
        *l3 = e - *l1 - *l2;
    }
}

int exec(const Params& p, long long* launches, char* err, int errcap, int* cuda)
{
    if (p.e < kMinLog2 || p.e > kMaxLog2) return failf(err, errcap, cuda, 0, "smfft: multi-pass transforms cover 2^15 .. 2^24 points%s", "");
    if (p.n_ffts <= 0) return 0;
    int dev = -1;
    BIG_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return failf(err, errcap, cuda, 0, "smfft: device ordinal out of range%s", "");
    DevState& st = g_state[dev];
    int l3, l2, l1;
    split(p.e, &l3, &l2, &l1);
    const bool three = l3 > 0;
    const long long N = 1LL << p.e, N1 = 1LL << l1, N2 = 1LL << l2, N3 = 1LL << l3;
    const PassInfo k1 = pass_for(three ? l3 : l2, p.dir, 0), k2 = pass_for(l2, p.dir, 0), k3 = pass_for(l1, p.dir, 1);
    float2 *wt_n = nullptr, *wt_23 = nullptr;
    {
        std::lock_guard<std::mutex> lk(st.mu);
        if (twiddle_table(st, p.e, &wt_n, err, errcap, cuda)) return 1;
        if (three && twiddle_table(st, l2 + l3, &wt_23, err, errcap, cuda)) return 1;
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
        for (const PassInfo* k : {&k1, &k2, &k3}) {
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
    auto launch = [&](const PassInfo& k, PassArgs& a, long long tiles, const char* what) {
        void* params[] = {&a};
        cudaError_t e = cudaLaunchKernel((const void*)k.fn, dim3((unsigned)tiles), dim3((unsigned)k.threads), params, (size_t)k.smem, p.stream);
        if (e != cudaSuccess) return failf(err, errcap, cuda, 1, "smfft: multi-pass transform, kernel launch: %s", cudaGetErrorString(e));
        (void)what;
        if (launches) *launches += 1;
        return 0;
    };
    auto box = [](long long len) { return len > 256 ? 256 : (int)len; };
    for (long long f0 = 0; f0 < p.n_ffts && !rc; f0 += chunk) {
        const long long cf = p.n_ffts - f0 < chunk ? p.n_ffts - f0 : chunk;
        const float2* in = (const float2*)p.in + f0 * N;
        float2* out = (float2*)p.out + f0 * N;
        PassArgs a;
        int bad = 0;
        if (three) {
            // pass 1: x -> scratch, columns of the [N3 cf rows][N1 N2 points] matrix
            memset(&a, 0, sizeof(a));
            a.base_tw = (const float2*)p.base_tw;
            a.wt = wt_23;
            bad |= host::encode_strided_map(&a.in_map, in, 2 * N1 * N2, N3 * cf, box(N3), 2 * k1.width);
            bad |= host::encode_strided_map(&a.out_map, scratch, 2 * N1 * N2, N3 * cf, box(N3), 2 * k1.width);
            a.groups = (int)(N1 * N2 / k1.width);
            a.col_shift = l1;  // column n1 + N1 n2 -> n2
            a.kdiv = 1;
            a.kscale = 1;
            if (bad) { rc = failf(err, errcap, cuda, 1, "smfft: cuTensorMapEncodeTiled failed (multi-pass transform, pass 1)%s", ""); break; }
            if ((rc = launch(k1, a, cf * a.groups, "pass 1"))) break;
        }
        // column pass over n2: two passes: x -> scratch; three passes: scratch -> scratch in place (a tile reads and writes the same box)
        memset(&a, 0, sizeof(a));
        a.base_tw = (const float2*)p.base_tw;
        a.wt = wt_n;
        bad |= host::encode_strided_map(&a.in_map, three ? (const void*)scratch : (const void*)in, 2 * N1, N2 * N3 * cf, box(N2), 2 * k2.width);
        bad |= host::encode_strided_map(&a.out_map, scratch, 2 * N1, N2 * N3 * cf, box(N2), 2 * k2.width);
        a.groups = (int)(N1 / k2.width);
        a.col_shift = 0;
        a.kdiv = (int)N3;    // slab = fft N3 + k3
        a.kscale = (int)N3;  // W_N^(n1 (k3 + N3 k2))
        if (bad) { rc = failf(err, errcap, cuda, 1, "smfft: cuTensorMapEncodeTiled failed (multi-pass transform, column pass)%s", ""); break; }
        if ((rc = launch(k2, a, cf * N3 * a.groups, "column pass"))) break;
        // row pass over n1: scratch -> X
        memset(&a, 0, sizeof(a));
        a.base_tw = (const float2*)p.base_tw;
        bad |= host::encode_tile_map(&a.in_map, scratch, cf * N / 16, three ? (int)(N1 / 16) : box(k3.width * N1 / 16));
        bad |= host::encode_strided_map(&a.out_map, out, 2 * N2 * N3, N1 * cf, box(N1), 2 * k3.width);
        a.groups = (int)((three ? N3 : N2) / k3.width);
        a.nmid = three ? (int)N2 : 1;
        a.out_col_scale = (int)N3;  // X[k3 + N3 k2 + N2 N3 k1]: the box of (k2, WD consecutive k3) starts at column N3 k2 + WD g
        a.kdiv = 1;
        if (bad) { rc = failf(err, errcap, cuda, 1, "smfft: cuTensorMapEncodeTiled failed (multi-pass transform, row pass)%s", ""); break; }
        if ((rc = launch(k3, a, cf * a.nmid * a.groups, "row pass"))) break;
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
