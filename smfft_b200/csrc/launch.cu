// launch.cu -- host launchers and the C ABI of libsmfft (include/smfft.h).
//
// Replaces the reference's host launchers
//   FFT_init / FFT_external_benchmark / FFT_multiple_benchmark    CT/FFT-GPU-32bit.cu:576-752
//   (same names)                                                   ST/...:299-384, RC/...:388-467
//   GPU_smFFT_4elements / GPU_smFFT_R2C / GPU_smFFT_C2R            CT:827-908, RC:572-688
// Differences by design: error codes instead of exit(1) (utils_cuda.h:12-22), 64-bit counts,
// persistent grids sized from the SM count instead of one CTA per FFT (CT:586-595).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "registry.hpp"
#include "tmap.hpp"
#include "smfft.h"

namespace smfft {
namespace host {

// ---------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::mutex g_mu;
static int g_opt_io = 0;  // 0 auto (measured best TMA staging per size), 1 LDG, 2 TMA in+out, 3 TMA in / registers out
static int g_opt_tw = TW_LUT;
static int g_opt_quirk4096 = 0;
static int g_opt_ctas_per_sm = 0;
static int g_opt_carveout = -2;  // -2: per kernel (see launch_batch), -1: driver default, 0..100: percent of shared memory
static cudaStream_t g_stream = 0;
static long long g_launches = 0;

struct DeviceState {
    int device = -1;
    int sms = 0;
    float2* tw = nullptr;
    std::vector<const void*> attr_done;
};
static std::vector<DeviceState> g_dev;

static int fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
    } while (0)

static int get_device_state(DeviceState** out)
{
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& d : g_dev)
        if (d.device == dev) { *out = &d; return 0; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail("smfft: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    DeviceState st;
    st.device = dev;
    st.sms = prop.multiProcessorCount;
    // twiddle table W_8192^j, forward sign, rounded from FP64 (twiddle.cuh)
    std::vector<float2> h(kTwiddleTableSize);
    for (int j = 0; j < kTwiddleTableSize; j++) {
        const double a = -2.0 * M_PI * (double)j / (double)kTwiddleTableSize;
        h[j] = make_float2((float)cos(a), (float)sin(a));
    }
    CUDA_TRY(cudaMalloc((void**)&st.tw, sizeof(float2) * kTwiddleTableSize));
    CUDA_TRY(cudaMemcpy(st.tw, h.data(), sizeof(float2) * kTwiddleTableSize, cudaMemcpyHostToDevice));
    if (!tensor_map_encoder()) return fail("smfft: cuTensorMapEncodeTiled not available in this driver");
    g_dev.push_back(st);
    *out = &g_dev.back();
    return 0;
}

// io: kernels::IO_* or -1 = the preferred TMA staging of this size and mode
static const KernelEntry* find_entry(int mode, int e, int dir, int reorder, int io, int tw, int reps)
{
    EntryList l{nullptr, 0};
    switch (e) {
        case 5: l = entries_e5(); break;
        case 6: l = entries_e6(); break;
        case 7: l = entries_e7(); break;
        case 8: l = entries_e8(); break;
        case 9: l = entries_e9(); break;
        case 10: l = entries_e10(); break;
        case 11: l = entries_e11(); break;
        case 12: l = entries_e12(); break;
        default: return nullptr;
    }
    for (int i = 0; i < l.count; i++) {
        const KernelEntry& k = l.entries[i];
        if (k.mode != mode || k.dir != dir || k.reorder != reorder || k.tw != tw || k.reps != reps) continue;
        if (io >= 0 ? k.io == io : k.prefer) return &k;
    }
    return nullptr;
}

static int make_map(CUtensorMap* m, const void* base, long long rows, int box_rows)
{
    const int r = encode_tile_map(m, base, rows, box_rows);
    if (r != 0) return fail("smfft: cuTensorMapEncodeTiled failed (CUresult %d, rows %lld, box %d)", r, rows, box_rows);
    return 0;
}

static int ilog2_exact(int n)
{
    for (int e = 0; e < 31; e++)
        if ((1 << e) == n) return e;
    return -1;
}

// launch one batch: n_points complex points = whole transforms of 2^e points each
static int launch_batch(int mode, int e, int dir, int reorder, int reps, const void* d_in, void* d_out, long long n_points)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    if (n_points <= 0) return 0;
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail("smfft: device pointers must be 16-byte aligned");
    int io = reps > 1 ? kernels::IO_LDG
                      : g_opt_io == 0 ? -1 : g_opt_io == 1 ? kernels::IO_LDG : g_opt_io == 2 ? kernels::IO_TMA
                      : g_opt_io == 3 ? kernels::IO_TMA_STG : kernels::IO_REG;
    const KernelEntry* k = find_entry(mode, e, dir, reorder, io, g_opt_tw, reps);
    if (!k && io == kernels::IO_REG) k = find_entry(mode, e, dir, reorder, -1, g_opt_tw, reps);  // no register-direct instance: the default one
    if (!k && io == kernels::IO_TMA_STG) k = find_entry(mode, e, dir, reorder, kernels::IO_TMA, g_opt_tw, reps);
    if (k && kernels::io_uses_tma(k->io) && n_points < k->tile_points)  // batch smaller than one tile: thread staging, no tensor map
        k = find_entry(mode, e, dir, reorder, kernels::IO_LDG, g_opt_tw, reps);
    if (k) io = k->io;
    if (!k) return fail("smfft: no kernel instance for mode %d, 2^%d points, dir %d, reorder %d, io %d, tw %d, reps %d", mode, e, dir, reorder, io, g_opt_tw, reps);

    bool need_attr = true;
    for (const void* f : ds->attr_done) need_attr &= (f != k->func);
    if (need_attr) {
        CUDA_TRY(cudaFuncSetAttribute(k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, k->smem_bytes));
        // TMA kernels: max shared (their loads bypass L1, a smaller carve-out only costs occupancy).  Register-direct kernels:
        // the driver's default, which leaves the large L1 their LDG/STG traffic needs ("cuFFT's 1.22 ms")
        const int carve = k->io == kernels::IO_REG ? (int)cudaSharedmemCarveoutDefault : (int)cudaSharedmemCarveoutMaxShared;
        CUDA_TRY(cudaFuncSetAttribute(k->func, cudaFuncAttributePreferredSharedMemoryCarveout, g_opt_carveout >= -1 ? g_opt_carveout : carve));
        ds->attr_done.push_back(k->func);
    }
    kernels::TileArgs args;
    memset(&args, 0, sizeof(args));
    args.n_points = n_points;
    args.n_tiles = (n_points + k->tile_points - 1) / k->tile_points;
    args.gin = (const float2*)d_in;
    args.gout = (float2*)d_out;
    args.tw = ds->tw;
    if (kernels::io_uses_tma(io)) {
        if (make_map(&args.in_map, d_in, n_points / 16, k->tile_points / 16)) return 1;
        if (io == kernels::IO_TMA && make_map(&args.out_map, d_out, n_points / 16, k->tile_points / 16)) return 1;
    }
    // persistent grid: SMs x CTAs/SM.  The TMA kernels are launched with the measured load concurrency
    // (tuning.hpp), never more than fits; everything else fills the SM.
    int fit = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, k->func, k->threads, k->smem_bytes));
    if (fit <= 0) return fail("smfft: kernel does not fit on an SM (smem %d B, %d threads)", k->smem_bytes, k->threads);
    int per_sm = g_opt_ctas_per_sm > 0 ? g_opt_ctas_per_sm : (k->ctas > 0 ? k->ctas : fit);
    if (per_sm > fit) per_sm = fit;
    long long grid = (long long)ds->sms * per_sm;
    if (k->ctas < 0 && g_opt_ctas_per_sm <= 0) grid = args.n_tiles;  // one CTA per tile
    if (grid > args.n_tiles) grid = args.n_tiles;
    if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;  // the kernels stride over the tiles, any grid is correct
    void* params[] = {&args};
    CUDA_TRY(cudaLaunchKernel(k->func, dim3((unsigned)grid), dim3((unsigned)k->threads), params, (size_t)k->smem_bytes, g_stream));
    g_launches++;
    return 0;
}

static int timed(double* ms, int (*fn)(void*), void* ctx)
{
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    CUDA_TRY(cudaEventRecord(a, g_stream));
    int rc = fn(ctx);
    cudaError_t e1 = cudaEventRecord(b, g_stream);
    cudaError_t e2 = cudaEventSynchronize(b);
    float t = 0.f;
    if (rc == 0 && e1 == cudaSuccess && e2 == cudaSuccess) {
        cudaEventElapsedTime(&t, a, b);
        if (ms) *ms += (double)t;  // accumulated, as CT:662
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (rc) return rc;
    if (e1 != cudaSuccess) return fail("cudaEventRecord: %s", cudaGetErrorString(e1));
    if (e2 != cudaSuccess) return fail("kernel execution failed: %s", cudaGetErrorString(e2));
    return 0;
}

struct Call {
    int mode, e, dir, reorder, reps;
    const void* in;
    void* out;
    long long n_points;
};
static int run_call(void* p)
{
    Call* c = (Call*)p;
    return launch_batch(c->mode, c->e, c->dir, c->reorder, c->reps, c->in, c->out, c->n_points);
}

static int c2c_call(Call* c, const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, int reps)
{
    const int e = ilog2_exact(fft_size);
    if (e < 5 || e > 12) return fail("smfft: wrong FFT length %d (C2C supports 32..4096)", fft_size);
    if (n_ffts < 0) return fail("smfft: negative nFFTs");
    int dir = inverse ? 1 : 0;
    if (g_opt_quirk4096 && fft_size == 4096 && inverse && !reorder) dir = 0;  // CT/SM_FFT_parameters.cuh:388
    long long ffts = n_ffts;
    if (reps > 1) ffts = n_ffts / reps;  // FFT_multiple: nFFTs/100 transforms' worth of data, 100 reps each (CT:669)
    *c = Call{kernels::MODE_C2C, e, dir, reorder ? 1 : 0, reps, d_in, d_out, ffts * fft_size};
    return 0;
}

static int r2c_call(Call* c, const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reps)
{
    const int en = ilog2_exact(fft_size);
    if (en < 6 || en > 13) return fail("smfft: wrong FFT length %d (R2C/C2R supports real 64..8192)", fft_size);
    long long ffts = reps > 1 ? n_ffts / reps : n_ffts;
    *c = Call{inverse ? kernels::MODE_C2R : kernels::MODE_R2C, en - 1, inverse ? 1 : 0, 1, reps, d_in, d_out, ffts * (fft_size / 2)};
    return 0;
}

}  // namespace host
}  // namespace smfft

using namespace smfft::host;

extern "C" {
#pragma GCC visibility push(default)

int smfft_version(void) { return SMFFT_VERSION; }
const char* smfft_last_error(void) { return g_err; }
long long smfft_launch_count(void) { return g_launches; }

int smfft_init(void)
{
    DeviceState* ds = nullptr;
    return get_device_state(&ds);
}

int smfft_set_stream(void* stream)
{
    g_stream = (cudaStream_t)stream;
    return 0;
}

int smfft_set_option(const char* key, int value)
{
    if (!strcmp(key, "io")) { if (value < 0 || value > 4) return fail("io must be 0..4"); g_opt_io = value; return 0; }
    if (!strcmp(key, "twiddle")) { if (value < 0 || value > 1) return fail("twiddle must be 0 or 1"); g_opt_tw = value; return 0; }
    if (!strcmp(key, "quirk_4096")) { g_opt_quirk4096 = value ? 1 : 0; return 0; }
    if (!strcmp(key, "ctas_per_sm")) { g_opt_ctas_per_sm = value; return 0; }
    if (!strcmp(key, "carveout")) {  // experiment switch: takes effect for kernels not launched yet (or after a new process)
        if (value < -2 || value > 100) return fail("carveout must be -2 (per kernel), -1 (driver default) or 0..100");
        g_opt_carveout = value;
        DeviceState* ds = nullptr;
        if (!get_device_state(&ds)) ds->attr_done.clear();
        return 0;
    }
    return fail("smfft: unknown option '%s'", key);
}

int smfft_get_option(const char* key)
{
    if (!strcmp(key, "io")) return g_opt_io;
    if (!strcmp(key, "twiddle")) return g_opt_tw;
    if (!strcmp(key, "quirk_4096")) return g_opt_quirk4096;
    if (!strcmp(key, "ctas_per_sm")) return g_opt_ctas_per_sm;
    if (!strcmp(key, "carveout")) return g_opt_carveout;
    if (!strcmp(key, "device_sms")) { DeviceState* ds = nullptr; return get_device_state(&ds) ? -1 : ds->sms; }
    return -1;
}

int smfft_exec_c2c(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder)
{
    Call c;
    if (c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 1)) return 1;
    return run_call(&c);
}

int smfft_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, double* ms)
{
    Call c;
    if (c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 1)) return 1;
    return timed(ms, run_call, &c);
}

int smfft_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, double* ms)
{
    if (n_ffts / 100 == 0) {  // CT:669-673
        if (ms) *ms = -1;
        return fail("smfft: FFT_multiple needs nFFTs >= 100");
    }
    Call c;
    if (c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 100)) return 1;
    return timed(ms, run_call, &c);
}

int smfft_stockham_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, double* ms)
{
    return smfft_external_benchmark(d_in, d_out, fft_size, n_ffts, inverse, 1, ms);
}

int smfft_stockham_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, double* ms)
{
    return smfft_multiple_benchmark(d_in, d_out, fft_size, n_ffts, inverse, 1, ms);
}

int smfft_exec_r2c_c2r(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse)
{
    Call c;
    if (r2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, 1)) return 1;
    return run_call(&c);
}

int smfft_r2c_c2r_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, double* ms)
{
    Call c;
    if (r2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, 1)) return 1;
    return timed(ms, run_call, &c);
}

int smfft_r2c_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, double* ms)
{
    if (n_ffts / 100 == 0) {
        if (ms) *ms = -1;
        return fail("smfft: FFT_multiple needs nFFTs >= 100");
    }
    Call c;
    if (r2c_call(&c, d_in, d_out, fft_size, n_ffts, 0, 100)) return 1;
    return timed(ms, run_call, &c);
}

// ---- host-pointer drivers ---------------------------------------------------------------------

static int host_driver(int is_real, const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder,
                       int n_runs, double* single_ms, double* multi_ms)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    if (n_runs < 1) n_runs = 1;
    const size_t bytes = (size_t)n_ffts * (size_t)fft_size * (is_real ? sizeof(float) : sizeof(float2));
    size_t free_mem = 0, total_mem = 0;
    CUDA_TRY(cudaMemGetInfo(&free_mem, &total_mem));
    if (2 * bytes > free_mem) return fail("smfft: not enough device memory (%zu B needed, %zu B free)", 2 * bytes, free_mem);  // CT:839-847
    void *d_in = nullptr, *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_in, bytes));
    if (cudaMalloc(&d_out, bytes) != cudaSuccess) { cudaFree(d_in); return fail("smfft: cudaMalloc failed"); }
    int rc = 0;
    double t_multi = 0, t_single = 0;
    if (multi_ms && n_ffts >= 100 && !(is_real && inverse)) {
        for (int r = 0; r < n_runs && !rc; r++) {
            if (cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, g_stream) != cudaSuccess) { rc = fail("H2D failed"); break; }
            rc = is_real ? smfft_r2c_multiple_benchmark(d_in, d_out, fft_size, n_ffts, &t_multi)
                         : smfft_multiple_benchmark(d_in, d_out, fft_size, n_ffts, inverse, reorder, &t_multi);
        }
        *multi_ms = t_multi / n_runs;
    }
    for (int r = 0; r < n_runs && !rc; r++) {
        if (cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, g_stream) != cudaSuccess) { rc = fail("H2D failed"); break; }
        rc = is_real ? smfft_r2c_c2r_external_benchmark(d_in, d_out, fft_size, n_ffts, inverse, &t_single)
                     : smfft_external_benchmark(d_in, d_out, fft_size, n_ffts, inverse, reorder, &t_single);
    }
    if (single_ms) *single_ms = t_single / n_runs;
    if (!rc && cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, g_stream) != cudaSuccess) rc = fail("D2H failed");
    if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fail("stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

int smfft_c2c_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder, int n_runs,
                   double* single_ms, double* multi_ms)
{
    return host_driver(0, h_in, h_out, fft_size, n_ffts, inverse, reorder, n_runs, single_ms, multi_ms);
}

int smfft_r2c_c2r_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int n_runs, double* single_ms,
                       double* multi_ms)
{
    return host_driver(1, h_in, h_out, fft_size, n_ffts, inverse, 1, n_runs, single_ms, multi_ms);
}

// Chunked pipeline: chunk i's H2D, FFT and D2H run on three streams chained by events, three device
// buffer pairs in rotation, so PCIe in both directions and the SMs overlap.  Buffers, streams and
// events are created once and kept (per process) so repeated calls pay no allocation.
namespace {
struct PipelineCtx {
    static const int NBUF = 3;
    size_t cap = 0;
    int device = -1;
    void* d_in[NBUF] = {nullptr, nullptr, nullptr};
    void* d_out[NBUF] = {nullptr, nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_fft = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[NBUF], ev_fft[NBUF], ev_out[NBUF], t0, t1;
    bool ready = false;
};
PipelineCtx g_pipe;

int pipeline_prepare(size_t chunk_bytes)
{
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    PipelineCtx& p = g_pipe;
    if (p.ready && p.device == dev && p.cap >= chunk_bytes) return 0;
    if (p.ready) {
        for (int i = 0; i < PipelineCtx::NBUF; i++) { cudaFree(p.d_in[i]); cudaFree(p.d_out[i]); p.d_in[i] = p.d_out[i] = nullptr; }
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&p.s_fft, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < PipelineCtx::NBUF; i++) {
            CUDA_TRY(cudaEventCreateWithFlags(&p.ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.ev_fft[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.ev_out[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreate(&p.t0));
        CUDA_TRY(cudaEventCreate(&p.t1));
    }
    p.ready = false;
    for (int i = 0; i < PipelineCtx::NBUF; i++) {
        CUDA_TRY(cudaMalloc(&p.d_in[i], chunk_bytes));
        CUDA_TRY(cudaMalloc(&p.d_out[i], chunk_bytes));
    }
    p.cap = chunk_bytes;
    p.device = dev;
    p.ready = true;
    return 0;
}
}  // namespace

int smfft_pipeline_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder, int mode,
                        long long chunk_ffts, double* ms)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    if (mode < 0 || mode > 2) return fail("smfft: pipeline mode must be 0 (C2C), 1 (R2C) or 2 (C2R)");
    const size_t fft_bytes = (size_t)fft_size * (mode == 0 ? sizeof(float2) : sizeof(float));
    if (chunk_ffts <= 0) chunk_ffts = (long long)((128u << 20) / fft_bytes);
    if (chunk_ffts > n_ffts) chunk_ffts = n_ffts;
    if (n_ffts <= 0) return 0;
    if (pipeline_prepare((size_t)chunk_ffts * fft_bytes)) return 1;
    PipelineCtx& p = g_pipe;
    const int NBUF = PipelineCtx::NBUF;
    int rc = 0;
    cudaStream_t saved = g_stream;
    CUDA_TRY(cudaEventRecord(p.t0, p.s_in));
    CUDA_TRY(cudaStreamWaitEvent(p.s_out, p.t0, 0));
    CUDA_TRY(cudaStreamWaitEvent(p.s_fft, p.t0, 0));
    long long done = 0;
    for (long long i = 0; done < n_ffts && !rc; i++, done += chunk_ffts) {
        const int b = (int)(i % NBUF);
        const long long cnt = (n_ffts - done) < chunk_ffts ? (n_ffts - done) : chunk_ffts;
        const char* src = (const char*)h_in + (size_t)done * fft_bytes;
        char* dst = (char*)h_out + (size_t)done * fft_bytes;
        if (i >= NBUF) {
            cudaStreamWaitEvent(p.s_in, p.ev_fft[b], 0);   // d_in[b] consumed by the FFT of chunk i-NBUF
            cudaStreamWaitEvent(p.s_fft, p.ev_out[b], 0);  // d_out[b] drained by the D2H of chunk i-NBUF
        }
        cudaMemcpyAsync(p.d_in[b], src, cnt * fft_bytes, cudaMemcpyHostToDevice, p.s_in);
        cudaEventRecord(p.ev_in[b], p.s_in);
        cudaStreamWaitEvent(p.s_fft, p.ev_in[b], 0);
        g_stream = p.s_fft;
        rc = mode == 0 ? smfft_exec_c2c(p.d_in[b], p.d_out[b], fft_size, cnt, inverse, reorder)
                       : smfft_exec_r2c_c2r(p.d_in[b], p.d_out[b], fft_size, cnt, mode == 2);
        g_stream = saved;
        cudaEventRecord(p.ev_fft[b], p.s_fft);
        cudaStreamWaitEvent(p.s_out, p.ev_fft[b], 0);
        cudaMemcpyAsync(dst, p.d_out[b], cnt * fft_bytes, cudaMemcpyDeviceToHost, p.s_out);
        cudaEventRecord(p.ev_out[b], p.s_out);
    }
    cudaEventRecord(p.t1, p.s_out);
    cudaError_t es = cudaEventSynchronize(p.t1);
    cudaStreamSynchronize(p.s_in);
    cudaStreamSynchronize(p.s_fft);
    if (!rc && es != cudaSuccess) rc = fail("smfft: pipeline failed: %s", cudaGetErrorString(es));
    if (!rc && ms) {
        float t = 0;
        cudaEventElapsedTime(&t, p.t0, p.t1);
        *ms += (double)t;
    }
    return rc;
}

}  // extern "C"
