/*
 * smfft.h -- C ABI of libsmfft (B200-native shared-memory batched FFT).
 *
 * Drop-in boundary for the SMFFT hot path.  The reference (KAdamek/SMFFT) has no FFI: its host
 * programs link C++-mangled launchers out of each .cu file.  Every entry point below names the
 * reference launcher it replaces (paths relative to the reference root):
 *   CT = SMFFT_CooleyTukey_C2C/FFT-GPU-32bit.cu
 *   ST = SMFFT_Stockham_C2C/FFT-GPU-32bit-Stockham.cu
 *   RC = SMFFT_Stockham_R2C_C2R/FFT-GPU-32bit-Stockham.cu
 * include/smfft_compat.hpp re-exports the same functions under the reference's exact C++
 * signatures so the reference's FFT.c objects link unmodified.
 *
 * Conventions (kept from the reference unless noted):
 *   - plain pointers and sizes only; device pointers are borrowed, out of place and must be
 *     16-byte aligned (cudaMalloc returns 256-byte alignment; anything else is rejected);
 *   - all transforms are un-normalised in both directions; C2R returns (N/2) * irfft;
 *   - `*ms` is ACCUMULATED (+=) with the elapsed milliseconds of the one launch, as in
 *     CT:662 / ST:343 / RC:431; the caller zeroes it;
 *   - return 0 = ok, non-zero = failure (smfft_last_error() has the text).  Unlike the reference
 *     (checkCudaErrors -> exit(1), utils_cuda.h:12-22) the library never terminates the process;
 *   - counts are 64-bit (the reference's int indexing stops at 4 GiB);
 *   - work is launched on the CUDA stream given to smfft_set_stream (default: legacy stream 0,
 *     as the reference) or passed to the *_stream entry points, on the current device.
 * Thread safety (the reference is single-threaded, one global device, CT:15): every entry point may be
 * called concurrently from several host threads, on one device or on several ("one host thread per
 * GPU").  The stream of smfft_set_stream, smfft_last_error() and smfft_last_error_code() are PER
 * THREAD; options (smfft_set_option) are process-wide; per-device state (twiddle table, kernel
 * attributes, pipeline buffers) is created on first use under a per-device lock;
 * smfft_pipeline_host calls on one device serialise.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef SMFFT_H
#define SMFFT_H

#ifdef __cplusplus
extern "C" {
#endif

#define SMFFT_VERSION 200

/* smfft_last_error_code(): what kind of failure the last non-zero return on this thread was.  The reference
 * tells them apart too: a wrong FFT length only prints (CT:656-658), a CUDA failure exits (utils_cuda.h:12-22). */
#define SMFFT_OK 0
#define SMFFT_ERR_ARGUMENT 1 /* wrong FFT length, bad count, misaligned pointer, unknown option */
#define SMFFT_ERR_CUDA 2     /* a CUDA runtime / driver call or the kernel itself failed */
#define SMFFT_ERR_MEMORY 3   /* not enough device memory (CT:839-847) */

/* replaces FFT_init() (CT:576-581, ST:299-303, RC:388-392): builds the twiddle table, opts the
 * kernels into >48 KB dynamic shared memory.  Idempotent; called lazily by everything else. */
int smfft_init(void);

/* ---- Cooley-Tukey C2C: N = 32..4096 (the reference's range), 8192 / 16384 (beyond it, external only) and
 * 2^15 .. 2^24 (two passes over HBM up to 2^20, three above, with a library-owned, stream-ordered scratch of min(batch, "two_pass_chunk_mib");
 * natural order only, in place allowed; the "io", "twiddle" and "select" options do not apply to these sizes;
 * smfft_pipeline_release() returns the scratch pool to the driver -- not while other calls run on that device) -----
 * replaces int FFT_external_benchmark(float2*, float2*, int FFT_size, int nFFTs, bool inverse,
 *                                     bool reorder, double* FFT_time)                  CT:583-664
 * reorder=1: natural-order DFT; reorder=0: DFT of the bit-reversed input.
 * d_in/d_out: float2[n_ffts * fft_size].  Any n_ffts >= 1 (no %4 / %2 rule for N = 32 / 64).
 * Unsupported fft_size returns an error (the reference prints and returns 0, CT:656-658). */
int smfft_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder,
                             double* ms);
/* replaces int FFT_multiple_benchmark(...same...)                                      CT:666-752
 * 100 in-place transforms per tile over n_ffts/100 FFTs' worth of tiles (timing only: values
 * overflow by design).  n_ffts < 100: *ms = -1, returns 1 (CT:670-673). */
int smfft_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder,
                             double* ms);
/* untimed launch of the same transform (what a caller embeds in its own stream): on the calling thread's
 * smfft_set_stream stream, or on an explicit cudaStream_t (as void*; NULL = legacy default stream) */
int smfft_exec_c2c(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder);
int smfft_exec_c2c_stream(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder,
                          void* stream);

/* ---- Stockham C2C: N = 32..2^24, natural order -------------------------------------------------
 * replaces void FFT_external_benchmark(float2*, float2*, int, int, double*)            ST:306-345
 *          void FFT_multiple_benchmark(float2*, float2*, int, int, double*)            ST:348-384
 * The reference's Stockham C2C directory is inverse-only; `inverse` selects the
 * direction here (the compat shim passes 1). */
int smfft_stockham_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse,
                                      double* ms);
int smfft_stockham_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse,
                                      double* ms);

/* ---- Stockham R2C / C2R: real N = 64..16384 (the reference: 512..4096; multiple: up to 8192) ----
 * replaces void FFT_external_benchmark(float*, float*, int FFT_size, int nFFTs, int inverse,
 *                                      double*)                                        RC:396-433
 *          void FFT_multiple_benchmark(float*, float*, int, int, double*)              RC:435-467
 * inverse=0 (R2C): d_in float[n_ffts*N] -> d_out float2[n_ffts*N/2], packed: bin0 = (X0.re, X_{N/2}.re)
 * inverse=1 (C2R): d_in packed float2[n_ffts*N/2] -> d_out float[n_ffts*N] = (N/2) * irfft. */
int smfft_r2c_c2r_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse,
                                     double* ms);
int smfft_r2c_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, double* ms);
int smfft_exec_r2c_c2r(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse);
int smfft_exec_r2c_c2r_stream(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, void* stream);

/* ---- the repeated path, untimed, with a repetition count that keeps the values finite -------------
 * SMFFT_DIT_multiple / FFT_GPU_multiple / FFT_GPU_R2C_C2R_multiple (CT:553-572, ST:262-278, RC:367-384) apply the
 * transform NREUSES = 100 times in place, so their outputs overflow and the reference never checks them.  This entry
 * runs the same kernels with reps = 3 (or 100) over ALL n_ffts transforms of d_in, so tests can compare F^3 x with
 * the oracle.  mode: 0 = C2C, 1 = R2C (forward). */
int smfft_exec_repeated(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, int mode,
                        int reps);

/* ---- host-pointer end-to-end drivers ------------------------------------------------------------
 * replaces int GPU_smFFT_4elements(float2* h_in, float2* h_out, int FFT_size, int nFFTs,
 *              bool inverse, bool reorder, int nRuns, double* single_ex_time,
 *              double* multi_ex_time)                                                  CT:827-908
 * Device memory check (CT:839-847), per-run H2D (CT:866,884), averaged times, D2H of the result.
 * Returns 1 on insufficient device memory, as the reference. */
int smfft_c2c_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder, int n_runs,
                   double* single_ms, double* multi_ms);
/* replaces int GPU_smFFT_R2C / GPU_smFFT_C2R (RC:572-688); inverse selects C2R */
int smfft_r2c_c2r_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int n_runs,
                       double* single_ms, double* multi_ms);
/* Chunked, double-buffered H2D -> FFT -> D2H pipeline on caller-provided (ideally pinned) host
 * buffers: the end-to-end path bench.py times as `e2e`.  mode: 0 = CT C2C, 1 = R2C, 2 = C2R. */
int smfft_pipeline_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder,
                        int mode, long long chunk_ffts, double* ms);
/* frees the current device's pipeline buffers (6 x chunk bytes), streams and events; they are re-created on the next
 * smfft_pipeline_host call */
int smfft_pipeline_release(void);

/* ---- knobs -------------------------------------------------------------------------------------
 * keys: "io" (0 = measured best TMA staging per size [default], 1 = LDG/STG staging by the threads,
 * 2 = TMA loads + TMA stores, 3 = TMA loads + stores from registers, 4 / 5 = register-direct input, shape A / B, where an
 * instance exists [natural-order C2C of 128..1024 points], the default elsewhere),
 * "select" (0 = the static, sustained-load table of tuning.hpp [default]; 1 = first-use selection: the first out-of-place
 * call of a transform on a batch of at least 2^"select_min_log2_points" [24] points times the table's instance against
 * its alternates on that batch -- synchronously, a few dozen launches -- and the fastest serves that transform on that
 * device from then on; smfft_select_report() lists the decisions, "select_reset" forgets them), "twiddle" (0 = table+powers
 * [default], 1 = MUFU __sincosf), "quirk_4096" (1 = reproduce FFT_4096_inverse_noreorder running
 * the forward transform, CT/SM_FFT_parameters.cuh:388; default 0 = mathematically correct),
 * "ctas_per_sm" (0 = built-in), "pipeline_chunk_mib" (default chunk of smfft_pipeline_host, 1..1024, default 128), "multi_pass_chunk_mib" (alias "two_pass_chunk_mib"; batch chunk = scratch size of the multi-pass transforms, 2^15 points and up, 1..65536, default 1024), "carveout" (experiment: -2 = per kernel [default], -1 = driver default, 0..100 = percent
 * of shared memory), "device_sms" (read-only). */
int smfft_set_option(const char* key, int value);
int smfft_get_option(const char* key);
/* the first-use selection's decisions on the current device as text, one line per transform with every candidate's
 * milliseconds; returns the number of bytes written into buf (NUL-terminated) */
int smfft_select_report(char* buf, int cap);
/* device address of the current device's twiddle table W_16384^j = exp(-2 pi i j / 16384), j = 0..16383 (float2, rounded
 * from FP64): what smfft::BlockFFT<..., TW_LUT>::fill_twiddles (include/smfft/device.cuh) reads.  NULL on failure. */
const void* smfft_twiddle_table(void);
/* CUDA stream (cudaStream_t as void*) used by THIS host thread's launches and event timing; NULL = legacy default */
int smfft_set_stream(void* stream);
/* number of kernels launched by this library since load (bench.py's gpu_launches evidence) */
long long smfft_launch_count(void);
const char* smfft_last_error(void);
int smfft_last_error_code(void);
int smfft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SMFFT_H */
