// kernels.cuh -- wrapper kernels of the native path: HBM -> shared tile -> block FFT -> HBM.
//
// Replaces the reference wrapper kernels
//   SMFFT_DIT_external / SMFFT_DIT_multiple        CT/FFT-GPU-32bit.cu:534-572
//   FFT_GPU_external / FFT_GPU_multiple            ST/FFT-GPU-32bit-Stockham.cu:243-278
//   FFT_GPU_R2C_C2R_external / _multiple           RC/FFT-GPU-32bit-Stockham.cu:349-384
// One kernel template covers them all: a tile of L = F*N contiguous points (F whole FFTs) is staged
// into shared memory in the SW128 layout, transformed in place by L/R threads, and written back.
//   IO_TMA : persistent CTAs, cp.async.bulk.tensor loads (mbarrier-tracked, STAGES buffers so the
//            load of tile k+1 and the store of tile k-1 overlap the FFT of tile k), TMA stores.
//   IO_LDG : same tiles staged by the threads with 16-byte LDG/STG (any 16-byte aligned pointer;
//            also the A/B comparison for the TMA path).
//   IO_REG : natural-order C2C / R2C without staging: each thread loads its own points from global memory
//            into registers (the next tile's while the current one is transformed, PF >= 0), shared memory
//            carries only the exchanges between passes, results leave from registers.
// REPS > 1 is the FFT_multiple benchmark: the transform is re-applied in place REPS times.
#pragma once
#include "smfft/detail/block_fft.cuh"
#include "smfft/detail/tma.cuh"

// FFT_multiple, natural order: chain the repetitions in registers (1) or go through the tile every time (0, for A/B)
#ifndef SMFFT_MULTIPLE_IN_REGISTERS
#define SMFFT_MULTIPLE_IN_REGISTERS 1
#endif

namespace smfft {
namespace kernels {

enum { IO_TMA = 0, IO_LDG = 1, IO_TMA_STG = 2, IO_REG = 3 };  // TMA_STG: TMA loads, results stored from registers
SMFFT_CX bool io_uses_tma(int io) { return io == IO_TMA || io == IO_TMA_STG; }
enum { MODE_C2C = detail::XF_C2C, MODE_R2C = detail::XF_R2C, MODE_C2R = detail::XF_C2R };

struct TileArgs {
    plat::TensorMap in_map;   // IO_TMA: [rows][32 x f32], box = tile rows, SWIZZLE_128B
    plat::TensorMap out_map;
    const float2* gin;        // IO_LDG
    float2* gout;
    long long n_tiles;
    long long n_points;       // valid float2 points in the batch (tail tile of IO_LDG)
    const float2* tw;         // W_16384^j table (forward sign)
    int l2_hint;              // bit 0: TMA loads evict_first, bit 1: TMA stores evict_first
};

// the per-tile transform, in place in shared memory; tile visible on entry, caller synchronises after
template <class C, int MODE, int REPS, class Hook = detail::NoHook>
SMFFT_DEV void tile_transform(float2* s, const float2* tw, Hook&& hook = Hook{})
{
    if constexpr (REPS == 1) {
        detail::block_fft_tile<C, MODE>(s, tw, hook);
    } else if constexpr (MODE == MODE_C2C && C::REORDER == 1 && !C::DUAL && SMFFT_MULTIPLE_IN_REGISTERS) {
        // Natural-order C2C ends every transform with the ownership it started with (v[m] = x[t + m T]), so the repetitions
        // chain in REGISTERS: the tile is read once and written once, shared memory carries only the exchanges between
        // passes -- what smfft::BlockFFT::exec (include/smfft/device.cuh) gives a user kernel that applies several
        // transforms in a row.  (fft_reorder = 0 re-reads the tile bit-reversed every time and keeps the loop below.)
        float2 v[C::R];
        const int tid = plat::tid();
        const int t = tid & (C::T - 1), fbase = (tid >> C::A) << C::E;
        detail::load_natural<C>(v, s, fbase, t);
#pragma unroll 1
        for (int rep = 0; rep < REPS; rep++) detail::run_passes<C, 0, detail::XF_C2C>(v, s, fbase, t, t, tw, detail::NoHook{});
        if constexpr ((!C::SAME_LAYOUT || !detail::LastExchangeSameAsTile<C>::value) && C::P > 1) plat::sync_block();
        detail::store_result<C, detail::XF_C2C>(v, s, fbase, t);
    } else {
#pragma unroll 1
        for (int rep = 0; rep < REPS; rep++) {
            detail::block_fft_tile<C, MODE>(s, tw);
            if (rep + 1 < REPS) plat::sync_block();
        }
    }
}

// same transform with the result stored straight from registers; hook: see detail::run_passes.
// REPS == 0 is the staging-only ceiling measurement (tools/tune).
template <class C, int MODE, int REPS, class Hook>
SMFFT_DEV void tile_transform_to_global(float2* s, const float2* tw, float2* __restrict__ g, long long valid,
                                        Hook&& hook)
{
    if constexpr (REPS == 0) {
        float2 v[C::R];
        const int tid = plat::tid();
        const int fbase = (tid >> C::A) << C::E, t = tid & (C::T - 1);
        detail::load_natural<C>(v, s, fbase, t);
        plat::sync_block();
        hook();
        detail::static_for<C::R>([&](auto M) {
            constexpr int m = decltype(M)::value;
            if (fbase + t + m * C::T < valid) plat::stg64_stream(g + fbase + t + m * C::T, v[m]);
        });
    } else {
        detail::block_fft_tile_to_global<C, MODE>(s, tw, g, valid, hook);
    }
}

template <class C>
SMFFT_DEV void coop_load_tile(float2* s, const float2* __restrict__ g, long long valid_points)
{
    constexpr int CHUNKS = C::L / 2, PER = CHUNKS / C::THREADS;
    static_assert(CHUNKS % C::THREADS == 0, "tile chunks must divide evenly over the CTA");
    const int tid = plat::tid();
    float4 q[PER];
    detail::static_for<PER>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const int ch = tid + i * C::THREADS;
        q[i] = (2LL * ch < valid_points) ? plat::ldg128_stream(g + 2 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    });
    detail::static_for<PER>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const int ch = tid + i * C::THREADS;
        plat::sts128(s + C::Layout::phys(2 * ch), q[i]);
    });
}

template <class C>
SMFFT_DEV void coop_store_tile(const float2* s, float2* __restrict__ g, long long valid_points)
{
    constexpr int CHUNKS = C::L / 2, PER = CHUNKS / C::THREADS;
    const int tid = plat::tid();
    detail::static_for<PER>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const int ch = tid + i * C::THREADS;
        const float4 q = plat::lds128(s + C::Layout::phys(2 * ch));
        if (2LL * ch < valid_points) plat::stg128_stream(g + 2 * ch, q);
    });
}

// PF: the pass after which the next tile's load is issued (-1: at the top of the iteration, IO_TMA only)
template <class C, int MODE, int IO, int STAGES, int REPS, int PF = (IO == IO_TMA ? -1 : 0)>
SMFFT_DEV void tile_kernel_body(const TileArgs& args, unsigned char* smem)
{
    constexpr int TILE_BYTES = C::L * 8;
    constexpr int ROWS = C::L / 16;  // 128-byte rows per tile
    constexpr int BOX_ROWS = ROWS > 256 ? 256 : ROWS;  // a TMA box holds at most 256 rows: larger tiles (8192 points) move as several boxes
    constexpr int NBOX = ROWS / BOX_ROWS;
    static_assert(TILE_BYTES % 1024 == 0, "tile must be a multiple of the 1 KB swizzle atom");
    static_assert(STAGES <= 8, "mbarrier block holds 8 barriers");
    const int tid = plat::tid();
    const long long first = plat::bid(), step = plat::nblocks();
    const long long my_tiles = first < args.n_tiles ? (args.n_tiles - first + step - 1) / step : 0;
    // compact twiddle table, once per (persistent) CTA; made visible by the first barrier below
    constexpr int NBUF = io_uses_tma(IO) ? STAGES : 1;
    float2* stw = reinterpret_cast<float2*>(smem + NBUF * TILE_BYTES + 64);
    // register-direct, one tile per CTA: the CTA lives for a few microseconds, so its own input loads go out FIRST and
    // the table fill (a global-memory round trip of its own) overlaps them instead of preceding them
    constexpr bool EARLY_LOAD = IO == IO_REG && PF < 0;
    // ... and where the plan has a pass behind which the table's stores can hide (C2C, at least two passes, few entries per
    // thread) the table's own global reads are issued next and parked in registers until the first exchange (TwiddlePrefetch)
    constexpr bool SPLIT_FILL = EARLY_LOAD && MODE == MODE_C2C && C::TW == TW_LUT && C::P >= 2 && detail::TwiddlePrefetch<C>::OK;
    float2 v0[EARLY_LOAD ? C::R : 1];
    detail::TwiddlePrefetch<C> twp;
    if constexpr (EARLY_LOAD) {
        if (my_tiles > 0) detail::load_global_natural<C>(v0, args.gin + first * C::L, args.n_points - first * C::L);
    }
    if constexpr (SPLIT_FILL)
        twp.load(args.tw, tid);
    else
        detail::fill_twiddle_table<C, MODE != MODE_C2C, MODE == MODE_C2R>(stw, args.tw, tid, C::THREADS);

    if constexpr (IO == IO_TMA || IO == IO_TMA_STG) {
        uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * TILE_BYTES);
        const uint64_t pol = args.l2_hint ? plat::l2_policy_evict_first() : 0;
        auto stage_ptr = [&](long long k) { return reinterpret_cast<float2*>(smem + (int)(k % STAGES) * TILE_BYTES); };
        auto issue_load = [&](long long k) {
            uint64_t* bar = &full[k % STAGES];
            plat::mbar_arrive_expect_tx(bar, TILE_BYTES);
            for (int b = 0; b < NBOX; b++) {
                float2* dst = stage_ptr(k) + b * (BOX_ROWS * 16);
                const int row = (int)((first + k * step) * ROWS) + b * BOX_ROWS;
                if (args.l2_hint & 1)
                    plat::tma_load_2d_hint(dst, &args.in_map, 0, row, bar, pol);
                else
                    plat::tma_load_2d(dst, &args.in_map, 0, row, bar);
            }
        };
        if (tid == 0) {
            for (int i = 0; i < STAGES; i++) plat::mbar_init(&full[i], 1);
            plat::mbar_fence_init();
            plat::tma_prefetch_desc(&args.in_map);
            if constexpr (IO == IO_TMA) plat::tma_prefetch_desc(&args.out_map);
        }
        plat::sync_block();
        if (tid == 0) {
            for (long long k = 0; k < STAGES - 1 && k < my_tiles; k++) issue_load(k);
        }
        if constexpr (IO == IO_TMA) {
            for (long long k = 0; k < my_tiles; k++) {
                // refill the buffer tile k-1 has just left, once its TMA store has finished reading it
                auto refill = [&]() {
                    const long long kn = k + STAGES - 1;
                    if (tid == 0 && kn < my_tiles) {
                        if (k > 0) plat::bulk_wait_read0();
                        issue_load(kn);
                    }
                };
                if constexpr (PF < 0 || REPS != 1) refill();
                plat::mbar_wait(&full[k % STAGES], (uint32_t)((k / STAGES) & 1));
                float2* s = stage_ptr(k);
                if constexpr (PF < 0 || REPS != 1)
                    tile_transform<C, MODE, REPS>(s, stw);
                else
                    tile_transform<C, MODE, REPS>(s, stw, detail::hook_at<PF>(refill));
                plat::fence_proxy_async();
                plat::sync_block();
                if (tid == 0) {
                    for (int b = 0; b < NBOX; b++) {
                        const int row = (int)((first + k * step) * ROWS) + b * BOX_ROWS;
                        if (args.l2_hint & 2)
                            plat::tma_store_2d_hint(&args.out_map, 0, row, s + b * (BOX_ROWS * 16), pol);
                        else
                            plat::tma_store_2d(&args.out_map, 0, row, s + b * (BOX_ROWS * 16));
                    }
                    plat::bulk_commit();
                }
            }
            if (tid == 0) plat::bulk_wait0();
        } else {
            // results leave from registers: a tile buffer is free as soon as every thread has passed
            // the first barrier of the NEXT tile; the refill is issued there or after a later pass (PF)
            if constexpr (STAGES == 1) {
                // ONE tile buffer (16384 points: 128 KB): the refill of the SAME buffer is issued as soon as the final exchange
                // has been read (hook_tail), so the next tile's load runs under the last pass and the stores from registers.
                // (Parking the lower half of the next tile in a spare 64 KB buffer earlier still changes nothing: the kernel
                // no longer waits for loads, profiles/r02_ab_16384_single_buffer.json.)
                static_assert(REPS == 1 && MODE == MODE_C2C && !C::DUAL, "single-buffer overlap: C2C external");
                if (tid == 0 && my_tiles > 0) issue_load(0);
                for (long long k = 0; k < my_tiles; k++) {
                    plat::mbar_wait(&full[0], (uint32_t)(k & 1));
                    const long long p0 = (first + k * step) * C::L;
                    auto refill = [&]() {
                        plat::fence_proxy_async();  // the exchanges wrote the buffer through the generic proxy
                        plat::sync_block();         // every thread holds its operands of the last pass in registers
                        if (tid == 0 && k + 1 < my_tiles) issue_load(k + 1);
                    };
                    tile_transform_to_global<C, MODE, REPS>(stage_ptr(0), stw, args.gout + p0, args.n_points - p0, detail::hook_tail(refill));
                }
            } else {
                for (long long k = 0; k < my_tiles; k++) {
                    plat::mbar_wait(&full[k % STAGES], (uint32_t)((k / STAGES) & 1));
                    const long long p0 = (first + k * step) * C::L;
                    auto refill = [&]() {
                        const long long kn = k + STAGES - 1;
                        if (tid == 0 && kn < my_tiles) issue_load(kn);
                    };
                    tile_transform_to_global<C, MODE, REPS>(stage_ptr(k), stw, args.gout + p0, args.n_points - p0,
                                                            detail::hook_at<(PF < 0 ? 0 : PF)>(refill));
                    // this buffer was written through the generic proxy (exchanges, real-pass scratch) and is refilled by a
                    // TMA load (async proxy) behind the first barrier of the next tile: order the two, as the PTX memory
                    // model asks (the IO_TMA path fences before its store for the same reason)
                    plat::fence_proxy_async();
                }
            }
        }
    } else if constexpr (IO == IO_REG) {
        static_assert(REPS == 1, "register-direct input is an external-benchmark path");
        float2* s = reinterpret_cast<float2*>(smem);
        if constexpr (!SPLIT_FILL) plat::sync_block();  // twiddle table
        auto tile_base = [&](long long k) { return (first + k * step) * C::L; };
        if constexpr (PF < 0) {
            for (long long k = 0; k < my_tiles; k++) {
                const long long p0 = tile_base(k);
                if (k > 0) detail::load_global_natural<C>(v0, args.gin + p0, args.n_points - p0);  // tile 0 was loaded before the table fill
                if constexpr (SPLIT_FILL) {
                    // the table lands in shared memory between the two barriers of the first exchange: pass 0 uses no
                    // twiddles, pass 1 reads them behind the second barrier (stored once: the first tile of the CTA)
                    auto park = [&]() { if (k == 0) twp.store(stw, tid); };
                    detail::block_fft_preloaded_to_global<C, MODE>(v0, s, stw, args.gout + p0, args.n_points - p0, detail::hook_at<0>(park));
                } else {
                    detail::block_fft_preloaded_to_global<C, MODE>(v0, s, stw, args.gout + p0, args.n_points - p0);
                }
            }
        } else {
            // software pipeline, unrolled by two so the register sets swap roles without moves
            float2 a[C::R], b[C::R];
            if (my_tiles > 0) detail::load_global_natural<C>(a, args.gin + tile_base(0), args.n_points - tile_base(0));
            for (long long k = 0; k < my_tiles; k += 2) {
                const long long p0 = tile_base(k), p1 = tile_base(k + 1), p2 = tile_base(k + 2);
                if (k + 1 < my_tiles) detail::load_global_natural<C>(b, args.gin + p1, args.n_points - p1);
                detail::block_fft_preloaded_to_global<C, MODE>(a, s, stw, args.gout + p0, args.n_points - p0);
                if (k + 1 < my_tiles) {
                    if (k + 2 < my_tiles) detail::load_global_natural<C>(a, args.gin + p2, args.n_points - p2);
                    detail::block_fft_preloaded_to_global<C, MODE>(b, s, stw, args.gout + p1, args.n_points - p1);
                }
            }
        }
    } else {
        float2* s = reinterpret_cast<float2*>(smem);
        plat::sync_block();
        for (long long k = 0; k < my_tiles; k++) {
            const long long p0 = (first + k * step) * C::L;
            const long long valid = args.n_points - p0;
            coop_load_tile<C>(s, args.gin + p0, valid);
            plat::sync_block();
            tile_transform<C, MODE, REPS>(s, stw);
            plat::sync_block();
            coop_store_tile<C>(s, args.gout + p0, valid);
            if (k + 1 < my_tiles) plat::sync_block();
        }
    }
}

// [tile buffers][64 B: mbarriers][twiddle table] + 1 KB alignment slack
template <class C, int IO, int STAGES, int MODE = MODE_C2C>
constexpr int smem_bytes()
{
    constexpr int tw = C::TW == TW_LUT ? (C::TW_C2C_ENTRIES + (MODE != MODE_C2C ? C::TW_R2C_ENTRIES : 0)) * 8 : 0;
    return (io_uses_tma(IO) ? STAGES : 1) * C::L * 8 + 64 + ((tw + 127) & ~127) + 1024;
}

#if !defined(SMFFT_EMU)
template <class C, int MODE, int IO, int STAGES, int REPS, int MINB, int PF = (IO == IO_TMA ? -1 : 0)>
__global__ void __launch_bounds__(C::THREADS, MINB) smfft_tile_kernel(const __grid_constant__ TileArgs args)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t a = plat::smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (a & 1023u)) & 1023u);  // SW128 needs a 1 KB aligned tile
    tile_kernel_body<C, MODE, IO, STAGES, REPS, PF>(args, smem);
}
#endif

}  // namespace kernels
}  // namespace smfft
