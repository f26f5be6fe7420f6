// smfft/detail/layout.cuh -- shared-memory layouts for a tile of float2 points.
//
// SW128 is the layout the whole native path lives in: 128-byte rows (16 float2), the 16-byte chunk
// index inside a row XORed with (row & 7).  It is exactly the TMA hardware SWIZZLE_128B pattern, so
// cp.async.bulk.tensor produces / consumes it directly, and it makes every access pattern of the
// FFT passes bank-conflict free at once:
//   * "column" accesses x = t + m*T by consecutive threads (64-bit): 16 lanes cover one row;
//   * "row" accesses of 16 contiguous points by consecutive threads (128-bit): 8 lanes hit 8 chunks;
//   * exchange writes j*r + q (first pass) and (j/Ns)*Ns*r + j%Ns + q*Ns, Ns >= 16 (later passes).
// The reference pads instead (stride 33 / 132 per warp, CT/FFT-GPU-32bit.cu:142-146, 252-256) and
// still has 2-way conflicts in reorder_256/512 and the Stockham writes.
#pragma once
#include "platform.cuh"

namespace smfft {
namespace detail {

struct LayoutSW128 {
    // x: logical float2 index inside the tile (tile base 1024-byte aligned)
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 4) & 7) << 1); }
};

// exchange layout for R = 32 points per thread: a thread's first-pass outputs are 32 contiguous points
// = two 128-byte rows, so the chunk index is XORed with (row pair & 7) to keep eight consecutive
// threads on eight different chunks
struct LayoutSW256 {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 5) & 7) << 1); }
};

// exchange layout for a scatter with Ns = 8 (reversed plans, small radix first): sixteen consecutive virtual threads write
// the same eight columns of two rows that are eight rows apart (same SW128 key); bit 3 of the row flips the upper chunk
// bit so the two groups land on different halves of the banks
struct LayoutSW128H {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ ((((x >> 4) & 7) ^ (((x >> 7) & 1) << 2)) << 1); }
};

// exchange layout after a radix-8 FIRST pass (reversed plans): 128-bit stores at x = 8j + q.  Eight consecutive virtual
// threads cover four rows (two chunks each); keying on (row & 3) keeps them apart AND makes the layout periodic in four
// rows, which is what the descending runs j' = 2T - t of the mirrored C2R ownership need: their wrapped lane lands on
// the chunk the missing lane of the previous group would have used (SW128, periodic in eight rows, collides 2-way there)
struct LayoutSW128Q {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 4) & 3) << 1); }
};

// reversed plans with a radix-4 first pass: 128-bit stores at x = 4j + q put eight consecutive virtual threads on the
// even chunks of two rows; keying on (row & 1) separates the rows and is periodic in two rows (the descending mirror runs
// need the periodicity, as in LayoutSW128Q)
struct LayoutSW128P {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 4) & 1) << 1); }
};
// ... and the Ns = 4 exchange that follows: sixteen consecutive virtual threads write the same four columns of four rows
// that are eight rows apart; bits 7..8 of x select the quarter of the row
struct LayoutSW128R4 {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 7) & 3) << 2); }
};

// exchange layout for FOUR points per thread (the reference-contract engines, 64-bit accesses only): the scatter after the
// first pass writes x = 4j + q and the one after the second x = 16 (j >> 2) + (j & 3) + 4q, so sixteen consecutive lanes
// cover four rows with the same four columns.  Folding the row's low two bits into BOTH column pairs separates them
// (low4' = ((a ^ b) << 2 | (q ^ b)) resp. ((a ^ b) | (q ^ b) << 2): a bijection of (a, b)); later scatters and every read
// cover whole rows, which any row-wise permutation keeps conflict-free.  (SW128 leaves these two scatters 2-way conflicted.)
struct LayoutSW4 {
    static SMFFT_HOST_DEV int phys(int x) { return x ^ (((x >> 4) & 3) * 5); }
};

struct LayoutLinear {
    static SMFFT_HOST_DEV int phys(int x) { return x; }
};

}  // namespace detail
}  // namespace smfft
