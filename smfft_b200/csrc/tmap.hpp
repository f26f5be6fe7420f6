// tmap.hpp -- host helper: TMA tensor map of a batch seen as rows of 128 bytes, SWIZZLE_128B.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace smfft {
namespace host {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
inline EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x 128 bytes, box = box_rows x 128 bytes; returns the CUresult (0 = ok), -1 without an encoder
// promo: 0 none, 1 64 B, 2 128 B, 3 256 B (L2 promotion);  swizzle: true = SWIZZLE_128B (LayoutSW128)
inline int encode_tile_map(CUtensorMap* m, const void* base, long long rows, int box_rows, int promo = 3, bool swizzle = true)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t gdim[2] = {32, (cuuint64_t)rows};  // 32 x f32 = one 128-byte row
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                               : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                            : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// a matrix of `rows` rows of row_f32 floats (row pitch = row_f32 * 4 bytes, a multiple of 128), box = box_rows x 128 bytes at any
// 128-byte column: the strided tiles of the two-pass transforms (big_fft.cu).  SWIZZLE_128B like encode_tile_map.
inline int encode_strided_map(CUtensorMap* m, const void* base, long long row_f32, long long rows, int box_rows, int box_f32 = 32)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t gdim[2] = {(cuuint64_t)row_f32, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)row_f32 * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_f32, (cuuint32_t)box_rows};  // 32 floats = one 128-byte line; 16 = 64-byte segments
    cuuint32_t estr[2] = {1, 1};
    return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_f32 == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,  // the swizzle span equals the box row: dense rows in shared memory
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace host
}  // namespace smfft
