// big_fft.hpp -- C2C transforms of 2^15 .. 2^24 points in two or three passes over HBM (host interface of big_fft.cu).
#pragma once
#include <cuda_runtime.h>

namespace smfft {
namespace big {

constexpr int kMinLog2 = 15, kMaxLog2 = 24;  // two passes up to 2^20 points, three from 2^21

struct Params {
    int e;                 // log2 of the transform length, kMinLog2 .. kMaxLog2
    int dir;               // 0 forward (exp -), 1 inverse (exp +), un-normalised
    const void* in;        // n_ffts transforms, contiguous, natural order
    void* out;             // same layout; may alias `in`
    long long n_ffts;
    cudaStream_t stream;
    const void* base_tw;   // the library's W_16384 table on this device (smfft_twiddle_table)
    long long chunk_bytes; // the batch is processed in chunks of about this many bytes (the intermediate of a chunk is scratch)
};

// 0 = ok; otherwise a message in err (cuda = 1 when a CUDA call failed, 2 when it was an allocation).  *launches: kernels launched.
int exec(const Params& p, long long* launches, char* err, int errcap, int* cuda);
// frees the per-device twiddle tables and the scratch pool of the current device
void release();

}  // namespace big
}  // namespace smfft
