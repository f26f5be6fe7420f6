// tune_shapes.cuh -- the shape list swept by tools/tune (measurement tool, not product)
#pragma once
#include <initializer_list>
#include <vector>

#include "registry.hpp"

using namespace smfft;
using namespace smfft::host;
using namespace smfft::kernels;

struct Variant {
    KernelEntry k;
    int b, tile_e;
    int hint = 0;        // TileArgs::l2_hint
    int out_off = 0;     // extra byte offset of the output buffer (DRAM bank-aliasing experiment)
    int promo = 3;       // tensor-map L2 promotion: 0 none, 1 64 B, 2 128 B, 3 256 B
    int swz = 1;         // tensor-map swizzle: 1 = 128B (product), 0 = none (staging-only experiments)
    int per_sm = 0;      // CTAs per SM to launch (0 = occupancy limit)
};

extern std::vector<Variant> g_variants;

template <int E, int B, int TILE_E, int STAGES, int MINB, int IO, int REPS = 1>
static void add_one(int hint = 0, int out_off = 0)
{
    if constexpr (TILE_E >= E && E > B && (TILE_E - B) <= 10 && (TILE_E - B) >= 5 && !(IO == IO_TMA_STG && STAGES < 2)) {
        g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 1, IO, TW_LUT, REPS>(), B, TILE_E, hint, out_off});
        if (REPS == 1 && hint == 0 && out_off == 0)
            g_variants.push_back(Variant{make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 0, IO, TW_LUT, REPS>(), B, TILE_E, 0, 0});
    }
}

// one shape, launched with an explicit number of CTAs per SM (grid = SMs x per_sm): what matters
// to the memory system is the requested load concurrency  per_sm x (STAGES-1) x tile bytes
template <int E, int B, int TILE_E, int STAGES, int MINB>
static void add_shape(std::initializer_list<int> per_sms)
{
    for (int per : per_sms) {
        const size_t before = g_variants.size();
        add_one<E, B, TILE_E, STAGES, MINB, IO_TMA>();
        add_one<E, B, TILE_E, STAGES, MINB, IO_TMA_STG>();
        for (size_t i = before; i < g_variants.size(); i++) g_variants[i].per_sm = per;
    }
}

template <int E>
static void add_size()
{
    add_shape<E, 4, 12, 2, 2>({2});        // first: reference output for the checks; 64 KB of loads requested
    add_shape<E, 4, 12, 2, 3>({3, 2});     // 96 / 64 KB
    add_shape<E, 4, 12, 3, 2>({2, 1});     // 128 / 64 KB
    add_shape<E, 4, 12, 1, 4>({4, 3});
    add_shape<E, 4, 11, 2, 4>({4, 3});     // 64 / 48 KB
    add_shape<E, 4, 11, 2, 6>({6, 5, 4, 3});
    add_shape<E, 4, 11, 3, 4>({2, 3});
    add_shape<E, 4, 10, 2, 8>({8, 6, 5});  // 64 / 48 / 40 KB
    add_shape<E, 3, 12, 2, 2>({2});        // R = 8: 512 threads per tile, 1024 threads at 64 KB
    add_shape<E, 3, 11, 2, 4>({4, 3});
    add_shape<E, 3, 11, 2, 6>({6, 4});
    add_shape<E, 3, 10, 2, 8>({8, 6});
    add_shape<E, 5, 12, 2, 2>({2});        // R = 32
    add_one<E, 4, 12, 1, 2, IO_LDG>();     // thread-staged comparison (no TMA)
    g_variants.push_back(Variant{make_entry_shape<E, 4, 12, 2, 2, MODE_C2C, 0, 1, IO_TMA, TW_MUFU, 1>(), 4, 12});
    g_variants.back().per_sm = 2;
    if constexpr (E == 10) {
        // staging-only ceilings (REPS = 0: tile in, tile out, no FFT)
        for (int per : {1, 2, 3}) {
            add_one<E, 4, 12, 2, 3, IO_TMA, 0>(); g_variants.back().per_sm = per;
            add_one<E, 4, 12, 2, 3, IO_TMA_STG, 0>(); g_variants.back().per_sm = per;
        }
        for (int per : {2, 3, 4, 6}) {
            add_one<E, 4, 11, 2, 6, IO_TMA, 0>(); g_variants.back().per_sm = per;
            add_one<E, 4, 11, 2, 6, IO_TMA_STG, 0>(); g_variants.back().per_sm = per;
        }
        add_one<E, 4, 12, 3, 2, IO_TMA, 0>(); g_variants.back().per_sm = 1;
    }
}
