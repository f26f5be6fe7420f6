// smfft/detail/tma.cuh -- TMA tensor copies + mbarrier, raw PTX for sm_100a.
//
// The HBM<->shared staging path of the native kernels (north_star subsystem 1): one elected thread
// issues cp.async.bulk.tensor (SASS UTMALDG / UTMASTG) for a whole tile; no thread spends registers
// or LSU issue slots on global memory.  The tensor map describes the batch as rows of 128 bytes with
// hardware SWIZZLE_128B, which is LayoutSW128 (layout.cuh).
// Replaces the reference's 4 x LDG.64 + 4 x STG.64 per thread staging (CT/FFT-GPU-32bit.cu:538-550).
#pragma once
#include "platform.cuh"

#if !defined(SMFFT_EMU)
#include <cuda.h>

namespace smfft {
namespace plat {

typedef CUtensorMap TensorMap;

SMFFT_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SMFFT_DEV void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SMFFT_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SMFFT_DEV void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SMFFT_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost TMA transaction traps (launch error) instead of hanging the device.
SMFFT_DEV void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}

SMFFT_DEV void tma_prefetch_desc(const TensorMap* m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global -> shared, 2-D tile, completes on an mbarrier with the byte count of the box
SMFFT_DEV void tma_load_2d(void* dst_smem, const TensorMap* m, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global, 2-D tile, tracked by the issuing thread's bulk async-group
SMFFT_DEV void tma_store_2d(const TensorMap* m, int c0, int c1, const void* src_smem)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(smem_u32(src_smem))
                 : "memory");
}
// L2 eviction-priority policies for streaming data (every byte of a batch is touched exactly once)
SMFFT_DEV uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
SMFFT_DEV void tma_load_2d_hint(void* dst_smem, const TensorMap* m, int c0, int c1, uint64_t* bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
SMFFT_DEV void tma_store_2d_hint(const TensorMap* m, int c0, int c1, const void* src_smem, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(smem_u32(src_smem)), "l"(pol)
                 : "memory");
}
SMFFT_DEV void stg64_hint(float2* p, float2 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
SMFFT_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have finished READING shared memory (buffers reusable)
SMFFT_DEV void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SMFFT_DEV void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// order this thread's generic-proxy shared-memory writes before later async-proxy (TMA) accesses
SMFFT_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace plat
}  // namespace smfft
#endif
