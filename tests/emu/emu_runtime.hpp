// emu_runtime.hpp -- host SIMT emulator for the smfft kernel sources.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the very same device templates (include/smfft/detail/*.cuh, smfft_b200/csrc/kernels.cuh)
// with g++ -DSMFFT_EMU and runs one CUDA block at a time, one ucontext fiber per CUDA thread:
//   * __syncthreads()            -> generation barrier between fibers
//   * mbarrier / TMA             -> asynchronous queue drained by the scheduler between rounds, so
//                                   a missing wait or a buffer reused too early shows up as wrong data
//   * shared-memory accessors    -> recorded per thread; after the block finishes the accesses are
//                                   grouped per warp instruction and bank wavefronts are counted
// Never linked into libsmfft.so; `pytest -m "not gpu"` uses it to check kernel logic without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

struct float2 {
    float x, y;
};
struct alignas(16) float4 {
    float x, y, z, w;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace smfft {
namespace emu {

struct SmemAccess {
    uint32_t addr;  // byte offset inside the block's shared memory
    uint8_t width;  // 8 or 16
    uint8_t store;
};

struct TmaOp {
    int kind;  // 0 load, 1 store
    unsigned char* smem;
    const struct TensorMapEmu* map;
    int row0;
    uint64_t* bar;
    int owner;
};

struct TensorMapEmu {
    unsigned char* base;  // global rows of 128 bytes
    long long total_rows;
    int box_rows;
};

struct BankStats {
    long long instr64 = 0, wave64 = 0, instr128 = 0, wave128 = 0;
    double factor() const
    {
        long long ideal = 2 * instr64 + 4 * instr128;
        return ideal ? (double)(wave64 + wave128) / (double)ideal : 1.0;
    }
};

struct Block {
    int nthreads = 0, cur = 0, bid = 0, nblocks = 1;
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<unsigned char>> stacks;
    std::vector<char> done;
    ucontext_t sched;
    int arrived = 0, alive = 0, gen = 0;
    // warp-level exchange state (shfl_xor / sync_warp): values parked per lane, a generation barrier per warp
    std::vector<float> shfl_val;
    std::vector<int> warp_arrived, warp_gen;
    long long shuffles = 0;
    unsigned char* smem = nullptr;
    size_t smem_bytes = 0;
    std::vector<TmaOp> queue;
    std::vector<std::vector<SmemAccess>> log;
    bool record = false;
    std::function<void()> body;
};

extern Block* g_blk;

inline void yield() { swapcontext(&g_blk->ctx[g_blk->cur], &g_blk->sched); }

void run_tma_op(const TmaOp& op);
void drain_tma(int owner_only, int kind_only);
void launch(int grid, int threads, size_t smem_bytes, const std::function<void(unsigned char*)>& body, BankStats* stats);

}  // namespace emu

namespace plat {

inline int tid() { return emu::g_blk->cur; }
inline int bid() { return emu::g_blk->bid; }
inline int nblocks() { return emu::g_blk->nblocks; }

inline void sync_block()
{
    emu::Block* b = emu::g_blk;
    const int gen = b->gen;
    if (++b->arrived == b->alive) {
        b->arrived = 0;
        b->gen++;
    } else {
        while (b->gen == gen) emu::yield();
    }
}

// all live lanes of the calling thread's warp (the kernels here never exit part of a warp early)
inline void sync_warp()
{
    emu::Block* b = emu::g_blk;
    const int w = b->cur / 32;
    const int lanes = std::min(32, b->nthreads - 32 * w);
    const int gen = b->warp_gen[w];
    if (++b->warp_arrived[w] == lanes) {
        b->warp_arrived[w] = 0;
        b->warp_gen[w]++;
    } else {
        while (b->warp_gen[w] == gen) emu::yield();
    }
}
inline float shfl_xor(float v, int mask)
{
    emu::Block* b = emu::g_blk;
    const int me = b->cur;
    b->shfl_val[me] = v;
    sync_warp();  // every lane has parked its value
    const int src = (me & ~31) | ((me ^ mask) & 31);
    const float r = src < b->nthreads ? b->shfl_val[src] : v;
    sync_warp();  // every lane has read before the slot is reused
    if ((me & 31) == 0) b->shuffles++;
    return r;
}

inline unsigned smem_addr(const void* p) { return (unsigned)((const unsigned char*)p - emu::g_blk->smem); }

inline void rec(const void* p, int width, int store)
{
    emu::Block* b = emu::g_blk;
    if (!b->record) return;
    const unsigned char* q = (const unsigned char*)p;
    if (q < b->smem || q >= b->smem + b->smem_bytes) {
        fprintf(stderr, "emu: shared-memory access out of bounds (tid %d)\n", b->cur);
        abort();
    }
    if (((uintptr_t)q) % width) {
        fprintf(stderr, "emu: misaligned %d-byte shared access (tid %d)\n", width, b->cur);
        abort();
    }
    b->log[b->cur].push_back(emu::SmemAccess{(uint32_t)(q - b->smem), (uint8_t)width, (uint8_t)store});
}

inline float2 lds64(const float2* p) { rec(p, 8, 0); return *p; }
inline void sts64(float2* p, float2 v) { rec(p, 8, 1); *p = v; }
inline float4 lds128(const float2* p) { rec(p, 16, 0); return *reinterpret_cast<const float4*>(p); }
inline void sts128(float2* p, float4 v) { rec(p, 16, 1); *reinterpret_cast<float4*>(p) = v; }

inline float2 ldg_ro(const float2* p) { return *p; }
inline void prefetch_l2(const void*) {}
inline float4 ldg128_stream(const float2* p) { float4 r; memcpy(&r, p, 16); return r; }
inline void stg128_stream(float2* p, float4 v) { memcpy(p, &v, 16); }
inline float2 ldg64_stream(const float2* p) { return *p; }
inline void stg64_stream(float2* p, float2 v) { *p = v; }
inline void fast_sincos(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
inline unsigned brev32(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}

// ---- TMA / mbarrier emulation ----
typedef emu::TensorMapEmu TensorMap;

inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }  // bit 0 = parity of the phase in progress, bits 32.. = bytes still expected
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { *bar += (uint64_t)bytes << 32; }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return (uint32_t)(*bar & 1u) != parity; }
inline void mbar_wait(uint64_t* bar, uint32_t parity)
{
    long spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > 1000000) { fprintf(stderr, "emu: mbarrier wait never satisfied (tid %d)\n", tid()); abort(); }
        emu::yield();
    }
}
inline void tma_prefetch_desc(const TensorMap*) {}
inline void tma_load_2d(void* dst, const TensorMap* m, int, int c1, uint64_t* bar)
{
    emu::g_blk->queue.push_back(emu::TmaOp{0, (unsigned char*)dst, m, c1, bar, tid()});
}
inline void tma_store_2d(const TensorMap* m, int, int c1, const void* src)
{
    emu::g_blk->queue.push_back(emu::TmaOp{1, (unsigned char*)src, m, c1, nullptr, tid()});
}
inline uint64_t l2_policy_evict_first() { return 0; }
inline void tma_load_2d_hint(void* dst, const TensorMap* m, int c0, int c1, uint64_t* bar, uint64_t) { tma_load_2d(dst, m, c0, c1, bar); }
inline void tma_store_2d_hint(const TensorMap* m, int c0, int c1, const void* src, uint64_t) { tma_store_2d(m, c0, c1, src); }
inline void stg64_hint(float2* p, float2 v, uint64_t) { *p = v; }
inline void bulk_commit() {}
inline void bulk_wait_read0() { emu::drain_tma(tid(), 1); }
inline void bulk_wait0() { emu::drain_tma(tid(), 1); }
inline void fence_proxy_async() {}

}  // namespace plat
}  // namespace smfft
