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
    if (op.kind == 0) {  // complete_tx: the phase flips when every expected byte has arrived (a tile may come as several boxes)
        const uint64_t bytes = (uint64_t)m->box_rows * 128;
        uint64_t pending = *op.bar >> 32;
        if (pending < bytes) { fprintf(stderr, "emu: TMA load completes more bytes than the mbarrier expects\n"); abort(); }
        pending -= bytes;
        *op.bar = (pending << 32) | (((*op.bar) & 1ull) ^ (pending == 0 ? 1ull : 0ull));
    }
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
        blk.shfl_val.assign(threads, 0.0f);
        blk.warp_arrived.assign((threads + 31) / 32, 0);
        blk.warp_gen.assign((threads + 31) / 32, 0);
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

#include "emu_run_cfg.hpp"

// what the product would run for (e, mode): arith flags | B << 4 | mirrored R2C << 8 | mirrored C2R << 9 | threads << 12
template <int E, int MODE>
static int product_flavour()
{
    using Tn = typename kernels::ShapeFor<E, MODE, 1, 1>::type;
    constexpr int ARITH = kernels::ArithFor<E, MODE, 1, 1>::value;
    using XL = typename std::conditional<Tn::B == 5, detail::LayoutSW256, detail::LayoutSW128>::type;
    using C = detail::BlockCfg<E, Tn::B, Tn::F, MODE == 2, 1, TW_LUT, detail::LayoutSW128, XL, true, true, ARITH>;
    return ARITH | (Tn::B << 4) | ((MODE == 1 && detail::MirrorR2C<C>::OK) << 8) | ((MODE == 2 && detail::MirrorC2R<C>::OK) << 9) | (C::THREADS << 12);
}

extern "C" {

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
        if (mode == 1 && reps == 1) return run_shape<E, Tr::B, Tr::F, 1, Tr::STAGES, 1, Tr::PF, kernels::ArithFor<E, 1, 1, 1>::value>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        using Tc = kernels::ShapeFor<E, 2, 1, 1>::type; /* C2R shape */                                              \
        if (mode == 2 && reps == 1) return run_shape<E, Tc::B, Tc::F, 2, Tc::STAGES, 1, Tc::PF, kernels::ArithFor<E, 2, 1, 1>::value>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        if (mode == 0 && reps == 1) return run_shape<E, Tn::B, Tn::F, 0, Tn::STAGES, 1, Tn::PF>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
        if (mode == 0 && reps == 3) return run_shape<E, Tn::B, Tn::F, 0, Tn::STAGES, 3, -1, kernels::ArithFor<E, 0, 1, 100>::value>(i, o, n_ffts, dir, reorder, io, tw, grid, bank_factor); \
    }
    SHAPE(5) SHAPE(6) SHAPE(7) SHAPE(8) SHAPE(9) SHAPE(10) SHAPE(11) SHAPE(12)
#undef SHAPE
    return -1;
}

int emu_product_flavour(int e, int mode)
{
#define PF(E) if (e == E) return mode == 0 ? product_flavour<E, 0>() : mode == 1 ? product_flavour<E, 1>() : product_flavour<E, 2>();
    PF(5) PF(6) PF(7) PF(8) PF(9) PF(10) PF(11) PF(12)
#undef PF
    return -1;
}

// the register-direct alternates of the tuning table: on | prefer << 1 | on_b << 2 | B << 4 | TILE_E << 8 | MINB << 16 (shape A), or shape B with which = 1
int emu_regdirect_info(int e, int which)
{
#define RD(E)                                                                                                                   \
    if (e == E) {                                                                                                               \
        using Rd = kernels::RegDirect<E>;                                                                                       \
        return (Rd::ON ? 1 : 0) | (Rd::PREFER ? 2 : 0) | (Rd::ON_B ? 4 : 0) | ((which ? Rd::B_B : Rd::B) << 4) |               \
               ((which ? Rd::TILE_E_B : Rd::TILE_E) << 8) | ((which ? Rd::MINB_B : Rd::MINB) << 16);                            \
    }
    RD(5) RD(6) RD(7) RD(8) RD(9) RD(10) RD(11) RD(12)
#undef RD
    return -1;
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
