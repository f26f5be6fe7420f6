// smfft/detail/compat_core.cuh -- which engine runs behind the reference-contract device functions.
//
// The contract (README.md:10-20 of the reference): the tile `s` is shared memory in natural, linear
// order, in place; blockDim.x == tile points / 4; the caller synchronises before and after.  Three engines:
//   * tiles held by ONE warp (32, 64, 128 points; four, two, one transform per 128-point tile): wf::warp_tile_fft --
//     registers and warp shuffles only between the one read and the one write of the tile, no block barrier;
//   * fft_reorder = 0 above 128 points: wf::block_fft_noreorder -- a 128-point warp-shuffle phase per warp, ONE exchange
//     through the tile, a second warp-shuffle phase for the remaining bits;
//   * natural order above 128 points: the Stockham radix-4 register passes of block_fft.cuh with swizzled exchanges.
// Platform-neutral (platform.cuh), so tests/emu runs exactly this dispatch on the CPU.
#pragma once
#include "block_fft.cuh"
#include "warp_fft.cuh"

namespace smfft {
namespace compat {

// the native block FFT under the reference's thread contract: R = 4 points per thread, linear
// tile on entry/exit, swizzled exchanges (LayoutSW4: conflict-free for every pass), 8-byte shared accesses only (no alignment demand)
template <int EXP, int FFTS_PER_TILE, int DIR, int REORDER>
using Cfg = detail::BlockCfg<EXP, 2, FFTS_PER_TILE, DIR, REORDER, TW_MUFU, detail::LayoutLinear, detail::LayoutSW4, false>;

#ifndef SMFFT_COMPAT_ENGINE
#define SMFFT_COMPAT_ENGINE 1  // 0: the shared-memory Stockham passes everywhere (the round-1 path, kept for A/B)
#endif

template <int EXP, int FFTS_PER_TILE, int DIR, int REORDER>
SMFFT_DEV void ct_dit(float2* s)
{
    static_assert(EXP >= 5 && EXP <= 12 && (FFTS_PER_TILE << EXP) == (EXP < 7 ? 128 : (1 << EXP)), "tile = max(128, N) points");
    if constexpr (SMFFT_COMPAT_ENGINE && EXP <= 7)
        detail::wf::warp_tile_fft<EXP, DIR, REORDER>(s);
    else if constexpr (SMFFT_COMPAT_ENGINE && !REORDER)
        detail::wf::block_fft_noreorder<EXP, DIR>(s);
    else
        detail::block_fft_tile<Cfg<EXP, FFTS_PER_TILE, DIR, REORDER>, detail::XF_C2C>(s, nullptr);
}

// The body of the external wrapper kernels (SMFFT_DIT_external<P>, CT:534-551): the tile at gin -> transform -> gout, with
// `s` (>= tile points) as scratch.  Same launch contract; the staging copies of the reference (4 LDG.64 -> 4 STS, barrier,
// ..., barrier, 4 LDS -> 4 STG.64) are folded into the transform's own first read and last write.
// The reference's launch shape (one CTA per tile, CT:586-595) caps the bytes in flight at (resident CTAs) x (tile bytes) per
// SM, so the larger tiles run at latency x concurrency, not at bandwidth.  There a CTA asks L2 for the tile a CTA launched
// SMFFT_COMPAT_PREFETCH_BYTES later will read: that CTA's loads then come from L2 instead of DRAM and its lifetime (= the
// concurrency each byte occupies) shrinks.  No extra DRAM traffic, no semantics.  Measured (profiles/
// r02_compat_prefetch_probe.json): 4096 points -5 % / -2 %, 256 and 1024 points fft_reorder = 0 -6 % / -3 %, natural-order
// 1024 points +3 % (off there).  The ONE-WARP tiles (N <= 128) do not react at all -- 2.17 ms with any distance, the
// reference 2.20 ms: 4.2 million one-warp CTAs are bound by the CTA launch rate (about one CTA per 77 ns and SM), which is
// why the external wrappers of those sizes sit at 1.01x the reference whatever the transform costs.
#ifndef SMFFT_COMPAT_PREFETCH_BYTES
#define SMFFT_COMPAT_PREFETCH_BYTES (8 << 20)
#endif
template <int TILE_POINTS>
SMFFT_DEV void prefetch_later_tile(const float2* __restrict__ gin)
{
    if constexpr (SMFFT_COMPAT_PREFETCH_BYTES > 0) {
        constexpr int TILE_BYTES = TILE_POINTS * 8, AHEAD = SMFFT_COMPAT_PREFETCH_BYTES / TILE_BYTES, LINES = TILE_BYTES / 128;
        const int t = plat::tid();
        if (t < LINES && plat::bid() + AHEAD < plat::nblocks())
            plat::prefetch_l2(reinterpret_cast<const char*>(gin) + (size_t)AHEAD * TILE_BYTES + (size_t)t * 128);
    }
}

template <int EXP, int FFTS_PER_TILE, int DIR, int REORDER>
SMFFT_DEV void ct_dit_external(float2* s, const float2* __restrict__ gin, float2* __restrict__ gout)
{
    if constexpr (EXP >= 11 || (EXP >= 8 && !REORDER)) prefetch_later_tile<(FFTS_PER_TILE << EXP)>(gin);
    if constexpr (SMFFT_COMPAT_ENGINE && EXP <= 7) {
        detail::wf::warp_tile_fft_global<EXP, DIR, REORDER>(gin, gout);
    } else if constexpr (SMFFT_COMPAT_ENGINE && !REORDER) {
        detail::wf::block_fft_noreorder_global<EXP, DIR>(s, gin, gout);
    } else if constexpr (SMFFT_COMPAT_ENGINE) {
        using C = Cfg<EXP, FFTS_PER_TILE, DIR, REORDER>;
        float2 v[C::R];
        detail::load_global_natural<C>(v, gin, C::L);
        detail::block_fft_preloaded_to_global<C, detail::XF_C2C>(v, s, nullptr, gout, C::L);
    } else {
        constexpr int L = FFTS_PER_TILE << EXP;
        const int t = plat::tid();
        detail::static_for<4>([&](auto Q) { plat::sts64(s + t + decltype(Q)::value * (L / 4), plat::ldg64_stream(gin + t + decltype(Q)::value * (L / 4))); });
        plat::sync_block();
        ct_dit<EXP, FFTS_PER_TILE, DIR, REORDER>(s);
        plat::sync_block();
        detail::static_for<4>([&](auto Q) { plat::stg64_stream(gout + t + decltype(Q)::value * (L / 4), plat::lds64(s + t + decltype(Q)::value * (L / 4))); });
    }
}

}  // namespace compat
}  // namespace smfft
