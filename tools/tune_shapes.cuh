// tune_shapes.cuh -- the shape list swept by tools/tune (measurement tool, not product)
#pragma once
#include <vector>

#include "registry.hpp"

using namespace smfft;
using namespace smfft::host;
using namespace smfft::kernels;

struct Variant {
    KernelEntry k;
    int b, tile_e;
};

extern std::vector<Variant> g_variants;

template <int E, int B, int TILE_E, int STAGES, int MINB>
static void add_shape()
{
    if constexpr (TILE_E >= E && E > B && (TILE_E - B) <= 10 && (TILE_E - B) >= 5) {
        for (int io = 0; io < 2; io++) {
            for (int ro = 1; ro >= 0; ro--) {
                KernelEntry k;
                if (io == 0 && ro == 1) k = make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 1, IO_TMA, TW_LUT, 1>();
                if (io == 0 && ro == 0) k = make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 0, IO_TMA, TW_LUT, 1>();
                if (io == 1 && ro == 1) k = make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 1, IO_LDG, TW_LUT, 1>();
                if (io == 1 && ro == 0) k = make_entry_shape<E, B, TILE_E, STAGES, MINB, MODE_C2C, 0, 0, IO_LDG, TW_LUT, 1>();
                if (io == 1 && STAGES != 2) continue;  // LDG path has no stages: time it once per shape
                g_variants.push_back(Variant{k, B, TILE_E});
            }
        }
    }
}

template <int E>
static void add_size()
{
    add_shape<E, 4, 11, 2, 4>();  // default first: reference output for the checks
    add_shape<E, 4, 11, 2, 6>();
    add_shape<E, 4, 11, 2, 8>();
    add_shape<E, 4, 11, 3, 4>();
    add_shape<E, 4, 11, 1, 6>();
    add_shape<E, 4, 12, 2, 3>();
    add_shape<E, 4, 12, 2, 2>();
    add_shape<E, 4, 12, 3, 2>();
    add_shape<E, 4, 12, 1, 4>();
    add_shape<E, 4, 13, 2, 1>();
    add_shape<E, 4, 10, 2, 8>();
    add_shape<E, 4, 10, 3, 8>();
    add_shape<E, 3, 11, 2, 4>();
    add_shape<E, 3, 11, 2, 6>();
    add_shape<E, 3, 10, 2, 8>();
    add_shape<E, 3, 12, 2, 2>();
    add_shape<E, 5, 12, 2, 2>();
    add_shape<E, 5, 12, 2, 3>();
    add_shape<E, 5, 13, 2, 1>();
    // MUFU twiddles on the default shape
    g_variants.push_back(Variant{make_entry_shape<E, 4, (E < 11 ? 11 : E), 2, (E == 12 ? 3 : 4), MODE_C2C, 0, 1, IO_TMA, TW_MUFU, 1>(), 4, (E < 11 ? 11 : E)});
}

