// smfft/detail/platform.cuh -- thin platform layer under the FFT core.
//
// Product build (nvcc, sm_100a): every function below is a one-line wrapper over a CUDA intrinsic.
// Test build (-DSMFFT_EMU, g++): tests/emu/emu_runtime.hpp supplies the same names on top of a
// host SIMT emulator (one fiber per CUDA thread), so the *same kernel source* is executed on the
// CPU by `pytest -m "not gpu"`.  The emulator is test infrastructure; it is never compiled into
// libsmfft.so and there is no CPU fallback in the product.
#pragma once

#if defined(SMFFT_EMU)
// emu_runtime.hpp must already be included by the test translation unit
#define SMFFT_DEV inline
#define SMFFT_HOST_DEV inline
#define SMFFT_CX constexpr
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define SMFFT_DEV __device__ __forceinline__
#define SMFFT_HOST_DEV __host__ __device__ __forceinline__
#define SMFFT_CX __host__ __device__ constexpr

namespace smfft {
namespace plat {

SMFFT_DEV int tid() { return (int)threadIdx.x; }
SMFFT_DEV int bid() { return (int)blockIdx.x; }
SMFFT_DEV int nblocks() { return (int)gridDim.x; }
SMFFT_DEV void sync_block() { __syncthreads(); }
SMFFT_DEV void sync_warp() { __syncwarp(); }
SMFFT_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// warp exchange, all 32 lanes (the reference's shfl / shfl_xor helpers, CT/FFT-GPU-32bit.cu:30-44)
SMFFT_DEV float shfl_xor(float v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// shared-memory accessors: plain dereferences (the compiler proves the shared address space after
// inlining and emits LDS/STS.64 and .128); kept as functions so the emulator can count bank conflicts.
SMFFT_DEV float2 lds64(const float2* p) { return *p; }
SMFFT_DEV void sts64(float2* p, float2 v) { *p = v; }
SMFFT_DEV float4 lds128(const float2* p) { return *reinterpret_cast<const float4*>(p); }
SMFFT_DEV void sts128(float2* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

SMFFT_DEV float2 ldg_ro(const float2* p) { return __ldg(p); }
// pull the 128-byte line at p into L2 (no register, no dependency): used by short-lived CTAs for the tile a LATER CTA will read
SMFFT_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
SMFFT_DEV float4 ldg128_stream(const float2* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
SMFFT_DEV void stg128_stream(float2* p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
// SMFFT_LDG64_VARIANT / SMFFT_STG64_VARIANT: cache-operator experiments of tools/tune (0 = product)
#ifndef SMFFT_LDG64_VARIANT
#define SMFFT_LDG64_VARIANT 0
#endif
#ifndef SMFFT_STG64_VARIANT
#define SMFFT_STG64_VARIANT 0
#endif
SMFFT_DEV float2 ldg64_stream(const float2* p)
{
#if SMFFT_LDG64_VARIANT == 1
    return *p;
#elif SMFFT_LDG64_VARIANT == 2
    return __ldcs(p);
#elif SMFFT_LDG64_VARIANT == 3
    return __ldcg(p);
#elif SMFFT_LDG64_VARIANT == 4
    return __ldg(p);
#else
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
#endif
}
SMFFT_DEV void stg64_stream(float2* p, float2 v)
{
#if SMFFT_STG64_VARIANT == 1
    *p = v;
#elif SMFFT_STG64_VARIANT == 2
    __stcs(p, v);
#elif SMFFT_STG64_VARIANT == 3
    __stcg(p, v);
#else
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
#endif
}

// MUFU sin/cos (the reference's --use_fast_math twiddle path, CT/FFT-GPU-32bit.cu:18-28)
SMFFT_DEV void fast_sincos(float a, float* s, float* c) { __sincosf(a, s, c); }
SMFFT_DEV unsigned brev32(unsigned v) { return __brev(v); }

}  // namespace plat
}  // namespace smfft
#endif
