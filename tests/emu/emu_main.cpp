// emu_main.cpp -- SIMT-emulator driver: runs the smfft kernel templates on the CPU.  TESTS ONLY.
// Built by tests/emu/build_emu.py into tests/emu/_build/libsmfft_emu.so and called through ctypes.
#include "emu_runtime.hpp"

#include "kernels.cuh"
#include "tuning.hpp"

namespace smfft {
namespace emu {

Block* g_blk = nullptr;

void run_tma_op(const TmaOp& op)
{
    const TensorMapEmu* m = op.map;
    for (int r = 0; r < m->box_rows; r++) {
        const long long grow = (long long)op.row0 + r;
        for (int c = 0; c < 8; c++) {
            unsigned char* sp = op.smem + (size_t)r * 128 + (size_t)((c ^ (r & 7)) * 16);  // SWIZZLE_128B
            if (op.kind == 0) {
                if (grow < m->total_rows)
                    memcpy(sp, m->base + grow * 128 + c * 16, 16);
                else
                    memset(sp, 0, 16);  // out-of-bounds rows are zero-filled
            } else if (grow < m->total_rows) {
                memcpy(m->base + grow * 128 + c * 16, sp, 16);  // out-of-bounds rows are clipped
            }
        }
    }
    if (op.kind == 0) *op.bar ^= 1ull;  // phase complete
}

void drain_tma(int owner_only, int kind_only)
{
    Block* b = g_blk;
    std::vector<TmaOp> keep;
    for (const TmaOp& op : b->queue) {
        if ((owner_only < 0 || op.owner == owner_only) && (kind_only < 0 || op.kind == kind_only))
            run_tma_op(op);
        else
            keep.push_back(op);
    }
    b->queue.swap(keep);
}

static void fiber_entry()
{
    Block* b = g_blk;
    b->body();
    b->done[b->cur] = 1;
    b->alive--;
    if (b->alive > 0 && b->arrived == b->alive) {  // exited threads release a pending barrier
        b->arrived = 0;
        b->gen++;
    }
    swapcontext(&b->ctx[b->cur], &b->sched);
}

static void analyse_banks(Block& b, BankStats* st)
{
    const int nw = (b.nthreads + 31) / 32;
    for (int w = 0; w < nw; w++) {
        const int l0 = w * 32, l1 = std::min(b.nthreads, l0 + 32);
        const size_t n = b.log[l0].size();
        bool uniform = true;
        for (int l = l0; l < l1; l++) uniform &= (b.log[l].size() == n);
        if (!uniform) continue;  // divergent warp (e.g. the bin-0 lane of the R2C pass): not analysed
        for (size_t i = 0; i < n; i++) {
            const int width = b.log[l0][i].width;
            const int group = width == 8 ? 16 : 8;  // lanes served together
            long long waves = 0;
            for (int g0 = l0; g0 < l1; g0 += group) {
                // per bank: set of distinct 32-bit words requested
                std::vector<uint32_t> words[32];
                for (int l = g0; l < std::min(l1, g0 + group); l++) {
                    const SmemAccess& a = b.log[l][i];
                    if (a.width != width) { fprintf(stderr, "emu: divergent access width\n"); abort(); }
                    for (int k = 0; k < width / 4; k++) {
                        const uint32_t word = a.addr / 4 + k;
                        auto& v = words[word & 31];
                        bool seen = false;
                        for (uint32_t x : v) seen |= (x == word);
                        if (!seen) v.push_back(word);
                    }
                }
                size_t deg = 1;
                for (int bk = 0; bk < 32; bk++) deg = std::max(deg, words[bk].size());
                waves += (long long)deg;
            }
            if (width == 8) { st->instr64++; st->wave64 += waves; }
            else { st->instr128++; st->wave128 += waves; }
        }
    }
}

void launch(int grid, int threads, size_t smem_bytes, const std::function<void(unsigned char*)>& body, BankStats* stats)
{
    const size_t STACK = 256 * 1024;
    for (int bid = 0; bid < grid; bid++) {
        Block blk;
        blk.nthreads = threads;
        blk.bid = bid;
        blk.nblocks = grid;
        blk.ctx.resize(threads);
        blk.stacks.assign(threads, std::vector<unsigned char>(STACK));
        blk.done.assign(threads, 0);
        blk.alive = threads;
        blk.log.resize(threads);
        blk.record = stats != nullptr && bid == 0;
        void* raw = nullptr;
        if (posix_memalign(&raw, 1024, smem_bytes + 1024)) abort();
        memset(raw, 0xCD, smem_bytes + 1024);
        blk.smem = (unsigned char*)raw;
        blk.smem_bytes = smem_bytes;
        blk.body = [&]() { body(blk.smem); };
        g_blk = &blk;
        for (int t = 0; t < threads; t++) {
            getcontext(&blk.ctx[t]);
            blk.ctx[t].uc_stack.ss_sp = blk.stacks[t].data();
            blk.ctx[t].uc_stack.ss_size = STACK;
            blk.ctx[t].uc_link = &blk.sched;
            makecontext(&blk.ctx[t], (void (*)())fiber_entry, 0);
        }
        long rounds = 0;
        while (blk.alive > 0) {
            for (int t = 0; t < threads; t++) {
                if (blk.done[t]) continue;
                blk.cur = t;
                swapcontext(&blk.sched, &blk.ctx[t]);
            }
            drain_tma(-1, -1);  // the async proxy makes progress between scheduling rounds
            if (++rounds > 10000000) { fprintf(stderr, "emu: block %d does not terminate\n", bid); abort(); }
        }
        drain_tma(-1, -1);
        if (blk.record) analyse_banks(blk, stats);
        free(raw);
        g_blk = nullptr;
    }
}

}  // namespace emu
}  // namespace smfft

// -------------------------------------------------------------------------------------------------
using namespace smfft;

static std::vector<float2> g_tw;
static const float2* twiddle_table()
{
    if (g_tw.empty()) {
        g_tw.resize(kTwiddleTableSize);
        for (int j = 0; j < kTwiddleTableSize; j++) {
            const double a = -2.0 * M_PI * (double)j / (double)kTwiddleTableSize;
            g_tw[j] = make_float2((float)cos(a), (float)sin(a));
        }
    }
    return g_tw.data();
}

template <int E, int B, int F, int MODE, int DIR, int REORDER, int IO, int TW, int STAGES, int REPS, int PF = (IO == kernels::IO_TMA ? -1 : 0), int DUAL = 0>
static int run_cfg(const float2* in, float2* out, long long n_ffts, int grid, double* bank_factor)
{
    using XL = typename std::conditional<B == 5, detail::LayoutSW256, detail::LayoutSW128>::type;
    using C = detail::BlockCfg<E, B, F, DIR, REORDER, TW, detail::LayoutSW128, XL, true, true, DUAL>;
    constexpr int ST = kernels::io_uses_tma(IO) ? STAGES : 1;
    kernels::TileArgs args;
    const long long n_points = n_ffts * C::N;
    const long long n_tiles = (n_points + C::L - 1) / C::L;
    args.in_map = emu::TensorMapEmu{(unsigned char*)in, n_points / 16, C::L / 16};
    args.out_map = emu::TensorMapEmu{(unsigned char*)out, n_points / 16, C::L / 16};
    args.gin = in;
    args.gout = out;
    args.n_tiles = n_tiles;
    args.n_points = n_points;
    args.tw = twiddle_table();
    args.l2_hint = 0;
    emu::BankStats st;
    if (grid <= 0 || grid > n_tiles) grid = (int)n_tiles;
    emu::launch(grid, C::THREADS, kernels::smem_bytes<C, IO, ST, MODE>(),
                [&](unsigned char* smem) { kernels::tile_kernel_body<C, MODE, IO, ST, REPS, PF>(args, smem); },
                bank_factor ? &st : nullptr);
    if (bank_factor) *bank_factor = st.factor();
    return 0;
}

// dispatch over (dir, reorder, io, tw) for a fixed shape
template <int E, int B, int F, int MODE, int STAGES, int REPS, int PFT = -1>
static int run_shape(const float2* in, float2* out, long long n_ffts, int dir, int reorder, int io, int tw, int grid,
                     double* bf)
{
#define CASE(D, RO, I, T)                                                     \
    if (dir == D && reorder == RO && io == I && tw == T)                      \
        return run_cfg<E, B, F, MODE, D, RO, I, T, STAGES, REPS, (I == kernels::IO_TMA ? PFT : (PFT < 0 ? 0 : PFT))>(in, out, n_ffts, grid, bf);
    if constexpr (MODE == kernels::MODE_C2C) {
        CASE(0, 1, 0, 0) CASE(0, 0, 0, 0) CASE(1, 1, 0, 0) CASE(1, 0, 0, 0)
        CASE(0, 1, 1, 0) CASE(0, 0, 1, 0) CASE(1, 1, 1, 0) CASE(1, 0, 1, 0)
        CASE(0, 1, 0, 1) CASE(0, 0, 0, 1) CASE(1, 1, 0, 1) CASE(1, 0, 0, 1)
        CASE(0, 1, 1, 1) CASE(1, 0, 1, 1)
        if constexpr (STAGES >= 2) { CASE(0, 1, 2, 0) CASE(0, 0, 2, 0) CASE(1, 1, 2, 0) CASE(1, 0, 2, 0) CASE(0, 1, 2, 1) }
    } else if constexpr (MODE == kernels::MODE_R2C) {
        CASE(0, 1, 0, 0) CASE(0, 1, 1, 0) CASE(0, 1, 0, 1)
        if constexpr (STAGES >= 2) { CASE(0, 1, 2, 0) }
    } else {
        CASE(1, 1, 0, 0) CASE(1, 1, 1, 0) CASE(1, 1, 0, 1)
        if constexpr (STAGES >= 2) { CASE(1, 1, 2, 0) }
    }
#undef CASE
    return -1;
}

// the configuration behind include/smfft/compat.cuh (reference thread contract: 4 points per thread,
// linear tile, swizzled exchanges, MUFU twiddles, 8-byte shared accesses)
template <int E, int MODE, int DIR, int REORDER>
static int run_compat(const float2* in, float2* out, long long n_ffts, double* bank_factor)
{
    using C = detail::BlockCfg<E, 2, (E < 7 ? (128 >> E) : 1), DIR, REORDER, TW_MUFU, detail::LayoutLinear, detail::LayoutSW128, false>;
    kernels::TileArgs args;
    const long long n_points = n_ffts * C::N;
    args.gin = in;
    args.gout = out;
    args.n_tiles = (n_points + C::L - 1) / C::L;
    args.n_points = n_points;
    args.tw = nullptr;
    args.l2_hint = 0;
    emu::BankStats st;
    emu::launch((int)args.n_tiles, C::THREADS, kernels::smem_bytes<C, kernels::IO_LDG, 1, MODE>(),
                [&](unsigned char* smem) { kernels::tile_kernel_body<C, MODE, kernels::IO_LDG, 1, 1>(args, smem); },
                bank_factor ? &st : nullptr);
    if (bank_factor) *bank_factor = st.factor();
    return 0;
}

// dual-lane transforms (block_fft_dual.cuh): two transforms per thread in the packed f32x2 lanes.
// kind: 0 C2C fwd natural, 1 C2C inv no-reorder, 2 C2C fwd no-reorder, 3 C2C inv natural, 4 R2C, 5 C2R; io as kernels::IO_*
template <int E, int B, int F, int STAGES, int PF>
static int run_dual_shape(const float2* i, float2* o, long long n, int kind, int io, int tw, int reps, int grid, double* bf)
{
    using namespace kernels;
#define DC(K, MODE, D, RO, I, T, RP)                                                                    \
    if (kind == K && io == I && tw == T && reps == RP)                                                  \
        return run_cfg<E, B, F, MODE, D, RO, I, T, STAGES, RP, (I == IO_TMA ? PF : (PF < 0 ? 0 : PF)), 1>(i, o, n, grid, bf);
    DC(0, 0, 0, 1, IO_TMA, 0, 1) DC(1, 0, 1, 0, IO_TMA, 0, 1) DC(2, 0, 0, 0, IO_TMA, 0, 1) DC(3, 0, 1, 1, IO_TMA, 0, 1)
    DC(4, 1, 0, 1, IO_TMA, 0, 1) DC(5, 2, 1, 1, IO_TMA, 0, 1)
    DC(0, 0, 0, 1, IO_TMA_STG, 0, 1) DC(1, 0, 1, 0, IO_TMA_STG, 0, 1) DC(4, 1, 0, 1, IO_TMA_STG, 0, 1) DC(5, 2, 1, 1, IO_TMA_STG, 0, 1)
    DC(0, 0, 0, 1, IO_LDG, 0, 1) DC(2, 0, 0, 0, IO_LDG, 0, 1) DC(4, 1, 0, 1, IO_LDG, 0, 1) DC(5, 2, 1, 1, IO_LDG, 0, 1)
    DC(0, 0, 0, 1, IO_TMA, 1, 1) DC(4, 1, 0, 1, IO_TMA, 1, 1) DC(5, 2, 1, 1, IO_LDG, 1, 1)
    DC(0, 0, 0, 1, IO_LDG, 0, 3) DC(2, 0, 0, 0, IO_LDG, 0, 3) DC(4, 1, 0, 1, IO_LDG, 0, 3)
#undef DC
    return -1;
}

extern "C" {

int emu_run_compat(const void* in, void* out, int e, long long n_ffts, int mode, int dir, int reorder, double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
#define CC(E)                                                                                      \
    if (e == E) {                                                                                  \
        if (mode == 0 && dir == 0 && reorder == 1) return run_compat<E, 0, 0, 1>(i, o, n_ffts, bank_factor); \
        if (mode == 0 && dir == 0 && reorder == 0) return run_compat<E, 0, 0, 0>(i, o, n_ffts, bank_factor); \
        if (mode == 0 && dir == 1 && reorder == 1) return run_compat<E, 0, 1, 1>(i, o, n_ffts, bank_factor); \
        if (mode == 0 && dir == 1 && reorder == 0) return run_compat<E, 0, 1, 0>(i, o, n_ffts, bank_factor); \
        if (mode == 1) return run_compat<E, 1, 0, 1>(i, o, n_ffts, bank_factor);                     \
        if (mode == 2) return run_compat<E, 2, 1, 1>(i, o, n_ffts, bank_factor);                     \
    }
    CC(5) CC(6) CC(7) CC(8) CC(9) CC(10) CC(11) CC(12)
#undef CC
    return -1;
}

// product shapes (tuning.hpp): e = log2 complex length
int emu_run(const void* in, void* out, int e, long long n_ffts, int mode, int dir, int reorder, int io, int tw,
            int reps, int grid, double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
#define SHAPE(E)                                                                                                    \
    if (e == E) {                                                                                                   \
        using Tn = kernels::Tuning<E>;                                                                              \
        using Tq = kernels::ShapeFor<E, 0, 1, 1>::type; /* natural-order shape (R = 32 for 512 / 1024) */           \
        if (mode == 0 && reps == 1 && reorder == 1) return run_shape<E, Tq::B, Tq::F, 0, Tq::STAGES, 1, Tq::PF>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        using Tr = kernels::ShapeFor<E, 1, 1, 1>::type; /* R2C / C2R shape */                                        \
        if (mode == 1 && reps == 1) return run_shape<E, Tr::B, Tr::F, 1, Tr::STAGES, 1, Tr::PF>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        if (mode == 2 && reps == 1) return run_shape<E, Tr::B, Tr::F, 2, Tr::STAGES, 1, Tr::PF>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        if (mode == 0 && reps == 1) return run_shape<E, Tn::B, Tn::F, 0, Tn::STAGES, 1, Tn::PF>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        if (mode == 0 && reps == 3) return run_shape<E, Tn::B, Tn::F, 0, Tn::STAGES, 3>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
    }
    SHAPE(5) SHAPE(6) SHAPE(7) SHAPE(8) SHAPE(9) SHAPE(10) SHAPE(11) SHAPE(12)
#undef SHAPE
    return -1;
}

// alternative shapes: exercise the generic pass machinery (other radices, tile sizes, stage counts)
int emu_run_alt(const void* in, void* out, int variant, long long n_ffts, int dir, int reorder, int io, int tw, int grid,
                double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    switch (variant) {
        case 0: return run_shape<10, 3, 1, 0, 1, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 1024 = 8*8*8*2, R=8
        case 1: return run_shape<10, 5, 4, 0, 3, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 1024 = 32*32, 3 stages
        case 2: return run_shape<9, 3, 4, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);    // 512 = 8*8*8
        case 3: return run_shape<7, 2, 8, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);    // 128, R=4 (compat shape)
        case 4: return run_shape<12, 4, 2, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 2 x 4096 per tile
        case 5: return run_shape<6, 3, 16, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 64 = 8*8
        case 6: return run_shape<5, 2, 16, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 32 = 4*4*2
        case 7: return run_shape<11, 5, 1, 0, 2, 1>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor);   // 2048 = 32*32*2
    }
    return -1;
}

// late prefetch (the refill of the previous buffer issued after pass PF instead of at the first barrier)
int emu_run_late(const void* in, void* out, int variant, long long n_ffts, int grid)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    switch (variant) {
        case 0: return run_cfg<12, 4, 1, 0, 0, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);
        case 1: return run_cfg<12, 4, 1, 0, 1, 0, kernels::IO_TMA_STG, TW_LUT, 2, 1, 2>(i, o, n_ffts, grid, nullptr);
        case 2: return run_cfg<10, 4, 4, 0, 0, 0, kernels::IO_TMA, TW_LUT, 2, 1, 2>(i, o, n_ffts, grid, nullptr);
        case 3: return run_cfg<8, 4, 8, 0, 0, 1, kernels::IO_TMA_STG, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);   // P = 2: clamps to pass 0
        case 4: return run_cfg<10, 4, 4, 1, 0, 1, kernels::IO_TMA_STG, TW_LUT, 3, 1, 1>(i, o, n_ffts, grid, nullptr);  // R2C, three buffers
        case 5: return run_cfg<11, 4, 2, 2, 1, 1, kernels::IO_TMA, TW_LUT, 2, 1, 1>(i, o, n_ffts, grid, nullptr);      // C2R
        // register-direct input (IO_REG): global -> registers, with and without the software prefetch
        case 6: return run_cfg<10, 4, 1, 0, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, 0>(i, o, n_ffts, grid, nullptr);
        case 7: return run_cfg<9, 5, 4, 0, 1, 1, kernels::IO_REG, TW_LUT, 1, 1, -1>(i, o, n_ffts, grid, nullptr);
        case 8: return run_cfg<10, 4, 2, 1, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, 0>(i, o, n_ffts, grid, nullptr);     // R2C
        case 9: return run_cfg<7, 4, 8, 1, 0, 1, kernels::IO_REG, TW_LUT, 1, 1, -1>(i, o, n_ffts, grid, nullptr);     // R2C
    }
    return -1;
}

int emu_run_dual(const void* in, void* out, int e, int b, long long n_ffts, int kind, int io, int tw, int reps, int grid,
                 double* bank_factor)
{
    const float2* i = (const float2*)in;
    float2* o = (float2*)out;
    if (b == 4) {
        switch (e) {
            case 8: return run_dual_shape<8, 4, 16, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 9: return run_dual_shape<9, 4, 8, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 10: return run_dual_shape<10, 4, 4, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 11: return run_dual_shape<11, 4, 2, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 12: return run_dual_shape<12, 4, 2, 2, 2>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
        }
    } else if (b == 5) {
        switch (e) {
            case 9: return run_dual_shape<9, 5, 8, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 10: return run_dual_shape<10, 5, 4, 2, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
            case 11: return run_dual_shape<11, 5, 2, 3, 1>(i, o, n_ffts, kind, io, tw, reps, grid, bank_factor);
        }
    }
    return -1;
}

int emu_alt_length(int variant)
{
    static const int n[] = {1024, 1024, 512, 128, 4096, 64, 32, 2048};
    return variant >= 0 && variant < 8 ? n[variant] : -1;
}

int emu_tile_points(int e)
{
    switch (e) {
#define TP(E) case E: return 1 << (kernels::TuningR32<E>::TILE_E > kernels::Tuning<E>::TILE_E ? kernels::TuningR32<E>::TILE_E : kernels::Tuning<E>::TILE_E);  // the largest product tile
        TP(5) TP(6) TP(7) TP(8) TP(9) TP(10) TP(11) TP(12)
#undef TP
    }
    return -1;
}
}
