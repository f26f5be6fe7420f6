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

#include <atomic>
#include <mutex>
#include <vector>

#include "big_fft.hpp"
#include "registry.hpp"
#include "tmap.hpp"
#include "smfft.h"

namespace smfft {
namespace host {

// ---------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------
// Thread-safety contract (include/smfft.h): every entry point may be called concurrently from several host threads,
// on the same or on different devices ("one host thread per GPU").  The error text, the error code and the
// stream set with smfft_set_stream are PER THREAD; options are process-wide atomics; per-device state lives in a fixed
// array indexed by the device ordinal (stable addresses) and is mutated only under that device's mutex.
static thread_local char g_err[512] = "";
static thread_local int g_err_code = SMFFT_OK;
static thread_local cudaStream_t t_stream = 0;  // smfft_set_stream: legacy default stream unless set, per host thread
static std::atomic<int> g_opt_io{0};  // 0 auto (measured best staging per size), 1 LDG, 2 TMA in+out, 3 TMA in / registers out, 4 register-direct
static std::atomic<int> g_opt_tw{TW_LUT};
static std::atomic<int> g_opt_quirk4096{0};
static std::atomic<int> g_opt_ctas_per_sm{0};
static std::atomic<int> g_opt_carveout{-2};  // -2: per kernel (see launch_batch), -1: driver default, 0..100: percent of shared memory
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_big_chunk_mib{1024};            // chunk of the two-pass transforms, 2^15 .. 2^18 points ("two_pass_chunk_mib")
static std::atomic<int> g_pipe_chunk_bytes{128 << 20};  // default chunk of smfft_pipeline_host ("pipeline_chunk_mib")

struct PipelineCtx {
    static const int NBUF = 3;
    std::mutex mu;       // one pipeline call at a time per device (the buffers are shared)
    size_t cap = 0;
    bool handles = false;  // streams and events exist (created once per device, kept)
    void* d_in[NBUF] = {nullptr, nullptr, nullptr};
    void* d_out[NBUF] = {nullptr, nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_fft = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[NBUF] = {}, ev_fft[NBUF] = {}, ev_out[NBUF] = {}, t0 = nullptr, t1 = nullptr;
};

struct DeviceState {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    int sms = 0;
    float2* tw = nullptr;
    std::vector<const void*> attr_done;  // guarded by mu
    PipelineCtx pipe;
};
static const int kMaxDevices = 64;
static DeviceState g_dev[kMaxDevices];

static int fail_code(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_err_code = code;
    return 1;
}
#define fail(...) fail_code(SMFFT_ERR_ARGUMENT, __VA_ARGS__)

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail_code(e__ == cudaErrorMemoryAllocation ? SMFFT_ERR_MEMORY : SMFFT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
    } while (0)

static int get_device_state(DeviceState** out)
{
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail("smfft: device ordinal %d out of range", dev);
    DeviceState& st = g_dev[dev];
    std::lock_guard<std::mutex> lk(st.mu);
    if (!st.ready) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            return fail_code(SMFFT_ERR_CUDA, "smfft: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
        // twiddle table W_16384^j, forward sign, rounded from FP64 (twiddle.cuh)
        std::vector<float2> h(kTwiddleTableSize);
        for (int j = 0; j < kTwiddleTableSize; j++) {
            const double a = -2.0 * M_PI * (double)j / (double)kTwiddleTableSize;
            h[j] = make_float2((float)cos(a), (float)sin(a));
        }
        if (!tensor_map_encoder()) return fail_code(SMFFT_ERR_CUDA, "smfft: cuTensorMapEncodeTiled not available in this driver");
        float2* tw = nullptr;
        CUDA_TRY(cudaMalloc((void**)&tw, sizeof(float2) * kTwiddleTableSize));
        cudaError_t e = cudaMemcpy(tw, h.data(), sizeof(float2) * kTwiddleTableSize, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(tw);
            return fail_code(SMFFT_ERR_CUDA, "smfft: twiddle table upload failed: %s", cudaGetErrorString(e));
        }
        st.device = dev;
        st.sms = prop.multiProcessorCount;
        st.tw = tw;
        st.ready = true;
    }
    *out = &st;
    return 0;
}

static EntryList entries_of(int e)
{
    switch (e) {
        case 5: return entries_e5();
        case 6: return entries_e6();
        case 7: return entries_e7();
        case 8: return entries_e8();
        case 9: return entries_e9();
        case 10: return entries_e10();
        case 11: return entries_e11();
        case 12: return entries_e12();
        case 13: return entries_e13();
        case 14: return entries_e14();
        default: return EntryList{nullptr, 0};
    }
}

// io: kernels::IO_* or -1 = the preferred staging of this size and mode; variant: 0 = the static table's instance,
// 1 / 2 = alternates (register-direct shapes A / B, alternate TMA shapes)
static const KernelEntry* find_entry(int mode, int e, int dir, int reorder, int io, int tw, int reps, int variant = 0)
{
    const EntryList l = entries_of(e);
    for (int i = 0; i < l.count; i++) {
        const KernelEntry& k = l.entries[i];
        if (k.mode != mode || k.dir != dir || k.reorder != reorder || k.tw != tw || k.reps != reps) continue;
        if (io >= 0 ? (k.io == io && k.variant == variant) : (k.prefer != 0)) return &k;
    }
    return nullptr;
}

// the alternates of one transform: candidates of the first-use selection (besides the static table's instance)
static int find_alternates(int mode, int e, int dir, int reorder, int tw, int reps, const KernelEntry** out, int cap)
{
    const EntryList l = entries_of(e);
    int n = 0;
    for (int i = 0; i < l.count && n < cap; i++) {
        const KernelEntry& k = l.entries[i];
        if (k.mode != mode || k.dir != dir || k.reorder != reorder || k.tw != tw || k.reps != reps) continue;
        if (!k.prefer && (k.variant >= 1 || k.tma_best)) out[n++] = &k;
    }
    return n;
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

// launch one kernel instance over the batch on `stream`
static int launch_entry(DeviceState* ds, const KernelEntry* k, const void* d_in, void* d_out, long long n_points, cudaStream_t stream)
{
    const int opt_ctas = g_opt_ctas_per_sm.load(), opt_carve = g_opt_carveout.load();
    {
        std::lock_guard<std::mutex> lk(ds->mu);
        bool need_attr = true;
        for (const void* f : ds->attr_done) need_attr &= (f != k->func);
        if (need_attr) {
            CUDA_TRY(cudaFuncSetAttribute(k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, k->smem_bytes));
            // TMA kernels: max shared (their loads bypass L1, a smaller carve-out only costs occupancy).  Register-direct kernels:
            // the driver's default, which leaves the large L1 their LDG/STG traffic needs ("cuFFT's 1.22 ms")
            const int carve = k->io == kernels::IO_REG ? (int)cudaSharedmemCarveoutDefault : (int)cudaSharedmemCarveoutMaxShared;
            CUDA_TRY(cudaFuncSetAttribute(k->func, cudaFuncAttributePreferredSharedMemoryCarveout, opt_carve >= -1 ? opt_carve : carve));
            ds->attr_done.push_back(k->func);
        }
    }
    kernels::TileArgs args;
    memset(&args, 0, sizeof(args));
    args.n_points = n_points;
    args.n_tiles = (n_points + k->tile_points - 1) / k->tile_points;
    args.gin = (const float2*)d_in;
    args.gout = (float2*)d_out;
    args.tw = ds->tw;
    if (kernels::io_uses_tma(k->io)) {
        const int box_rows = k->tile_points / 16 > 256 ? 256 : k->tile_points / 16;  // TMA boxes hold at most 256 rows (kernels.cuh splits larger tiles)
        if (make_map(&args.in_map, d_in, n_points / 16, box_rows)) return 1;
        if (k->io == kernels::IO_TMA && make_map(&args.out_map, d_out, n_points / 16, box_rows)) return 1;
    }
    // persistent grid: SMs x CTAs/SM.  The TMA kernels are launched with the measured load concurrency
    // (tuning.hpp), never more than fits; everything else fills the SM.
    int fit = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, k->func, k->threads, k->smem_bytes));
    if (fit <= 0) return fail_code(SMFFT_ERR_CUDA, "smfft: kernel does not fit on an SM (smem %d B, %d threads)", k->smem_bytes, k->threads);
    int per_sm = opt_ctas > 0 ? opt_ctas : (k->ctas > 0 ? k->ctas : fit);
    if (per_sm > fit) per_sm = fit;
    long long grid = (long long)ds->sms * per_sm;
    if (k->ctas < 0 && opt_ctas <= 0) grid = args.n_tiles;  // one CTA per tile
    if (grid > args.n_tiles) grid = args.n_tiles;
    if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;  // the kernels stride over the tiles, any grid is correct
    void* params[] = {&args};
    CUDA_TRY(cudaLaunchKernel(k->func, dim3((unsigned)grid), dim3((unsigned)k->threads), params, (size_t)k->smem_bytes, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ---- first-use selection (option "select" = 1) ----------------------------------------------------------------------
// Where single launches and a long, power-capped run disagree about the best shape (tuning.hpp: 4096 points R = 32 vs
// R = 16, register-direct vs TMA staging for 128..1024 points, one large vs six small CTAs at 32 points) no compile-time
// table is right for every box and caller.  With "select" = 1 the first sufficiently large out-of-place call of a
// transform times the static table's instance and its alternates ON THE CALLER'S OWN BATCH -- interleaved rounds of
// back-to-back launches between CUDA events -- and the fastest serves that transform on that device from then on.
// The call that tunes is synchronous and runs the transform a few dozen times (its output is the same every time).
struct TuneKey {
    int mode, e, dir, reorder, tw;
    bool operator==(const TuneKey& o) const { return mode == o.mode && e == o.e && dir == o.dir && reorder == o.reorder && tw == o.tw; }
};
struct TuneRecord {
    TuneKey key;
    const KernelEntry* chosen;
    int ncand;
    const KernelEntry* cand[4];
    float ms[4];
    long long n_points;
};
static std::mutex g_tune_mu;
static std::vector<TuneRecord> g_tuned[kMaxDevices];
static std::atomic<int> g_opt_select{0};               // 0 static table, 1 first-use selection
static std::atomic<int> g_opt_select_min_log2{24};     // smallest batch (log2 complex points) worth timing: 2^24 = 128 MiB

static const KernelEntry* tuned_lookup(int dev, const TuneKey& key)
{
    std::lock_guard<std::mutex> lk(g_tune_mu);
    for (const TuneRecord& r : g_tuned[dev])
        if (r.key == key) return r.chosen;
    return nullptr;
}

static int autotune(DeviceState* ds, const TuneKey& key, const KernelEntry* base, const void* d_in, void* d_out, long long n_points,
                    cudaStream_t stream, const KernelEntry** chosen)
{
    TuneRecord rec;
    memset(&rec, 0, sizeof(rec));
    rec.key = key;
    rec.n_points = n_points;
    rec.cand[0] = base;
    rec.ncand = 1 + find_alternates(key.mode, key.e, key.dir, key.reorder, key.tw, 1, rec.cand + 1, 3);
    *chosen = base;
    if (rec.ncand > 1) {
        const int ROUNDS = 3, BURST = 4;
        cudaEvent_t ev[4][3][2];
        bool ok = true;
        for (int c = 0; c < rec.ncand; c++)
            for (int r = 0; r < ROUNDS; r++)
                for (int j = 0; j < 2; j++) ok &= cudaEventCreate(&ev[c][r][j]) == cudaSuccess;
        int rc = launch_entry(ds, rec.cand[0], d_in, d_out, n_points, stream);  // warm-up: attributes, code, L2
        for (int c = 1; c < rec.ncand && !rc;) {  // an alternate that cannot be launched here (shared memory, registers) is dropped, not fatal
            if (launch_entry(ds, rec.cand[c], d_in, d_out, n_points, stream)) {
                for (int j = c; j + 1 < rec.ncand; j++) rec.cand[j] = rec.cand[j + 1];
                rec.ncand--;
                g_err[0] = 0;
                g_err_code = SMFFT_OK;
                cudaGetLastError();
            } else {
                c++;
            }
        }
        for (int r = 0; r < ROUNDS && !rc; r++)
            for (int c = 0; c < rec.ncand && !rc; c++) {
                ok &= cudaEventRecord(ev[c][r][0], stream) == cudaSuccess;
                for (int b = 0; b < BURST && !rc; b++) rc = launch_entry(ds, rec.cand[c], d_in, d_out, n_points, stream);
                ok &= cudaEventRecord(ev[c][r][1], stream) == cudaSuccess;
            }
        ok &= cudaStreamSynchronize(stream) == cudaSuccess;
        if (!rc && ok) {
            int best = 0;
            for (int c = 0; c < rec.ncand; c++) {
                float t[3] = {0, 0, 0};
                for (int r = 0; r < ROUNDS; r++) cudaEventElapsedTime(&t[r], ev[c][r][0], ev[c][r][1]);
                const float lo = fminf(t[0], fminf(t[1], t[2])), hi = fmaxf(t[0], fmaxf(t[1], t[2]));
                rec.ms[c] = (t[0] + t[1] + t[2] - lo - hi) / BURST;  // median round
                if (rec.ms[c] < rec.ms[best]) best = c;
            }
            // an alternate must win by more than the noise of the protocol to displace the table's instance
            if (best != 0 && rec.ms[best] > 0.995f * rec.ms[0]) best = 0;
            *chosen = rec.cand[best];
        }
        for (int c = 0; c < rec.ncand; c++)
            for (int r = 0; r < ROUNDS; r++)
                for (int j = 0; j < 2; j++) cudaEventDestroy(ev[c][r][j]);
        if (rc) return rc;
        if (!ok) return fail_code(SMFFT_ERR_CUDA, "smfft: first-use selection failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    rec.chosen = *chosen;
    std::lock_guard<std::mutex> lk(g_tune_mu);
    g_tuned[ds->device].push_back(rec);
    return 0;
}

// launch one batch on `stream`: n_points complex points = whole transforms of 2^e points each
static int launch_batch(int mode, int e, int dir, int reorder, int reps, const void* d_in, void* d_out, long long n_points,
                        cudaStream_t stream)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    if (n_points <= 0) return 0;
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail("smfft: device pointers must be 16-byte aligned");
    if (e >= big::kMinLog2) {
        // 2^15 .. 2^24 points: two or three passes over HBM (big_fft.cu), natural order, C2C
        big::Params bp{e, dir, d_in, d_out, n_points >> e, stream, ds->tw, (long long)g_big_chunk_mib.load() << 20};
        long long launched = 0;
        int cuda = 0;
        const int rc = big::exec(bp, &launched, g_err, (int)sizeof(g_err), &cuda);
        g_launches.fetch_add(launched, std::memory_order_relaxed);
        if (rc) g_err_code = cuda == 2 ? SMFFT_ERR_MEMORY : cuda == 1 ? SMFFT_ERR_CUDA : SMFFT_ERR_ARGUMENT;
        return rc;
    }
    const int opt_io = g_opt_io.load(), opt_tw = g_opt_tw.load();
    int io = reps > 1 ? kernels::IO_LDG
                      : opt_io == 0 ? -1 : opt_io == 1 ? kernels::IO_LDG : opt_io == 2 ? kernels::IO_TMA
                      : opt_io == 3 ? kernels::IO_TMA_STG : kernels::IO_REG;
    const int variant = io == kernels::IO_REG ? (opt_io == 5 ? 2 : 1) : 0;
    const KernelEntry* k = find_entry(mode, e, dir, reorder, io, opt_tw, reps, variant);
    if (!k && io == kernels::IO_REG) k = find_entry(mode, e, dir, reorder, -1, opt_tw, reps);  // no register-direct instance: the default one
    if (!k && io == kernels::IO_TMA_STG) k = find_entry(mode, e, dir, reorder, kernels::IO_TMA, opt_tw, reps);
    if (k && kernels::io_uses_tma(k->io) && n_points < k->tile_points)  // batch smaller than one tile: thread staging, no tensor map
        k = find_entry(mode, e, dir, reorder, kernels::IO_LDG, opt_tw, reps);
    if (!k) return fail("smfft: no kernel instance for mode %d, 2^%d points, dir %d, reorder %d, io %d, tw %d, reps %d", mode, e, dir, reorder, io, opt_tw, reps);

    if (g_opt_select.load() == 1 && opt_io == 0 && reps == 1 && n_points >= (1LL << g_opt_select_min_log2.load())) {
        const size_t bytes = (size_t)n_points * sizeof(float2);
        const char *a = (const char*)d_in, *b = (const char*)d_out;
        const bool overlap = a < b + bytes && b < a + bytes;   // in place: the transform cannot be repeated
        if (!overlap) {
            const TuneKey key{mode, e, dir, reorder, opt_tw};
            const KernelEntry* t = tuned_lookup(ds->device, key);
            if (!t && autotune(ds, key, k, d_in, d_out, n_points, stream, &t)) return 1;
            if (t) k = t;
        }
    }
    return launch_entry(ds, k, d_in, d_out, n_points, stream);
}

struct Call {
    int mode, e, dir, reorder, reps;
    const void* in;
    void* out;
    long long n_points;
};
static int run_call(const Call& c, cudaStream_t stream)
{
    return launch_batch(c.mode, c.e, c.dir, c.reorder, c.reps, c.in, c.out, c.n_points, stream);
}

// one launch between two events on `stream`; the elapsed milliseconds are ACCUMULATED into *ms (CT:662)
static int timed(double* ms, const Call& c, cudaStream_t stream)
{
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return fail_code(SMFFT_ERR_CUDA, "cudaEventCreate failed"); }
    cudaError_t e0 = cudaEventRecord(a, stream);
    int rc = e0 == cudaSuccess ? run_call(c, stream) : 0;
    cudaError_t e1 = cudaEventRecord(b, stream);
    cudaError_t e2 = cudaEventSynchronize(b);
    float t = 0.f;
    if (rc == 0 && e0 == cudaSuccess && e1 == cudaSuccess && e2 == cudaSuccess) {
        cudaEventElapsedTime(&t, a, b);
        if (ms) *ms += (double)t;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (rc) return rc;
    if (e0 != cudaSuccess) return fail_code(SMFFT_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e0));
    if (e1 != cudaSuccess) return fail_code(SMFFT_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e1));
    if (e2 != cudaSuccess) return fail_code(SMFFT_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e2));
    return 0;
}

static int c2c_call(Call* c, const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, int reps)
{
    const int e = ilog2_exact(fft_size);
    if (e < 5 || e > big::kMaxLog2 || (e >= 13 && reps > 1))
        return fail("smfft: wrong FFT length %d (C2C supports 32..16777216, FFT_multiple 32..4096)", fft_size);
    if (e >= big::kMinLog2 && !reorder) return fail("smfft: transforms of %d points (several passes over HBM) run in natural order only (reorder = 1)", fft_size);
    if (n_ffts < 0) return fail("smfft: negative nFFTs");
    int dir = inverse ? 1 : 0;
    if (g_opt_quirk4096.load() && fft_size == 4096 && inverse && !reorder) dir = 0;  // CT/SM_FFT_parameters.cuh:388
    long long ffts = n_ffts;
    if (reps > 1) ffts = n_ffts / reps;  // FFT_multiple: nFFTs/100 transforms' worth of data, 100 reps each (CT:669)
    *c = Call{kernels::MODE_C2C, e, dir, reorder ? 1 : 0, reps, d_in, d_out, ffts * fft_size};
    return 0;
}

static int r2c_call(Call* c, const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reps)
{
    const int en = ilog2_exact(fft_size);
    if (en < 6 || en > 14 || (en == 14 && reps > 1)) return fail("smfft: wrong FFT length %d (R2C/C2R supports real 64..16384, FFT_multiple up to 8192)", fft_size);
    if (n_ffts < 0) return fail("smfft: negative nFFTs");
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
int smfft_last_error_code(void) { return g_err_code; }
long long smfft_launch_count(void) { return g_launches.load(); }

int smfft_init(void)
{
    DeviceState* ds = nullptr;
    return get_device_state(&ds);
}

// what the first-use selection decided on the current device, one line per transform:
// "mode e dir reorder tw points | chosen io variant threads tile | io/variant:ms ..."; returns the number of bytes written
int smfft_select_report(char* buf, int cap)
{
    if (!buf || cap <= 0) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    std::lock_guard<std::mutex> lk(g_tune_mu);
    int n = 0;
    buf[0] = 0;
    for (const TuneRecord& r : g_tuned[dev]) {
        n += snprintf(buf + n, n < cap ? cap - n : 0, "mode %d e %d dir %d reorder %d tw %d points %lld | chosen io %d variant %d threads %d tile %d |",
                      r.key.mode, r.key.e, r.key.dir, r.key.reorder, r.key.tw, r.n_points, r.chosen->io, r.chosen->variant, r.chosen->threads, r.chosen->tile_points);
        for (int c = 0; c < r.ncand && n < cap; c++) n += snprintf(buf + n, cap - n, " %d/%d:%.4f", r.cand[c]->io, r.cand[c]->variant, r.ms[c]);
        if (n < cap) n += snprintf(buf + n, cap - n, "\n");
        if (n >= cap) { n = cap - 1; break; }
    }
    return n;
}

const void* smfft_twiddle_table(void)
{
    DeviceState* ds = nullptr;
    return get_device_state(&ds) ? nullptr : (const void*)ds->tw;
}

int smfft_set_stream(void* stream)
{
    t_stream = (cudaStream_t)stream;
    return 0;
}

int smfft_set_option(const char* key, int value)
{
    if (!key) return fail("smfft: null option key");
    if (!strcmp(key, "io")) { if (value < 0 || value > 5) return fail("io must be 0..5"); g_opt_io = value; return 0; }
    if (!strcmp(key, "select")) {
        if (value < 0 || value > 1) return fail("select must be 0 (static table) or 1 (first-use selection)");
        g_opt_select = value;
        return 0;
    }
    if (!strcmp(key, "select_min_log2_points")) { if (value < 10 || value > 40) return fail("select_min_log2_points must be 10..40"); g_opt_select_min_log2 = value; return 0; }
    if (!strcmp(key, "select_reset")) {  // forget what has been selected (every device)
        std::lock_guard<std::mutex> lk(g_tune_mu);
        for (auto& v : g_tuned) v.clear();
        return 0;
    }
    if (!strcmp(key, "twiddle")) { if (value < 0 || value > 1) return fail("twiddle must be 0 or 1"); g_opt_tw = value; return 0; }
    if (!strcmp(key, "quirk_4096")) { g_opt_quirk4096 = value ? 1 : 0; return 0; }
    if (!strcmp(key, "ctas_per_sm")) { g_opt_ctas_per_sm = value; return 0; }
    if (!strcmp(key, "two_pass_chunk_mib") || !strcmp(key, "multi_pass_chunk_mib")) {  // two names, one option: the passes became three from 2^21 points
        if (value < 1 || value > 65536) return fail("multi_pass_chunk_mib must be 1..65536");
        g_big_chunk_mib = value;
        return 0;
    }
    if (!strcmp(key, "pipeline_chunk_mib")) { if (value < 1 || value > 1024) return fail("pipeline_chunk_mib must be 1..1024"); g_pipe_chunk_bytes = value << 20; return 0; }
    if (!strcmp(key, "carveout")) {  // experiment switch: takes effect for kernels not launched yet (or after a new process)
        if (value < -2 || value > 100) return fail("carveout must be -2 (per kernel), -1 (driver default) or 0..100");
        g_opt_carveout = value;
        DeviceState* ds = nullptr;
        if (!get_device_state(&ds)) {
            std::lock_guard<std::mutex> lk(ds->mu);
            ds->attr_done.clear();
        }
        return 0;
    }
    return fail("smfft: unknown option '%s'", key);
}

int smfft_get_option(const char* key)
{
    if (!key) return -1;
    if (!strcmp(key, "io")) return g_opt_io;
    if (!strcmp(key, "select")) return g_opt_select;
    if (!strcmp(key, "select_min_log2_points")) return g_opt_select_min_log2;
    if (!strcmp(key, "twiddle")) return g_opt_tw;
    if (!strcmp(key, "quirk_4096")) return g_opt_quirk4096;
    if (!strcmp(key, "ctas_per_sm")) return g_opt_ctas_per_sm;
    if (!strcmp(key, "carveout")) return g_opt_carveout;
    if (!strcmp(key, "two_pass_chunk_mib") || !strcmp(key, "multi_pass_chunk_mib")) return g_big_chunk_mib;
    if (!strcmp(key, "pipeline_chunk_mib")) return g_pipe_chunk_bytes >> 20;
    if (!strcmp(key, "device_sms")) { DeviceState* ds = nullptr; return get_device_state(&ds) ? -1 : ds->sms; }
    return -1;
}

int smfft_exec_c2c_stream(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, void* stream)
{
    Call c;
    if (c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 1)) return 1;
    return run_call(c, (cudaStream_t)stream);
}

int smfft_exec_c2c(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder)
{
    return smfft_exec_c2c_stream(d_in, d_out, fft_size, n_ffts, inverse, reorder, t_stream);
}

int smfft_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, double* ms)
{
    Call c;
    if (c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 1)) return 1;
    return timed(ms, c, t_stream);
}

int smfft_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, double* ms)
{
    if (n_ffts / 100 == 0) {  // CT:669-673
        if (ms) *ms = -1;
        return fail("smfft: FFT_multiple needs nFFTs >= 100");
    }
    Call c;
    if (c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 100)) return 1;
    return timed(ms, c, t_stream);
}

int smfft_stockham_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, double* ms)
{
    return smfft_external_benchmark(d_in, d_out, fft_size, n_ffts, inverse, 1, ms);
}

int smfft_stockham_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, double* ms)
{
    return smfft_multiple_benchmark(d_in, d_out, fft_size, n_ffts, inverse, 1, ms);
}

int smfft_exec_r2c_c2r_stream(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, void* stream)
{
    Call c;
    if (r2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, 1)) return 1;
    return run_call(c, (cudaStream_t)stream);
}

int smfft_exec_r2c_c2r(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse)
{
    return smfft_exec_r2c_c2r_stream(d_in, d_out, fft_size, n_ffts, inverse, t_stream);
}

int smfft_r2c_c2r_external_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, double* ms)
{
    Call c;
    if (r2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, 1)) return 1;
    return timed(ms, c, t_stream);
}

int smfft_r2c_multiple_benchmark(const void* d_in, void* d_out, int fft_size, long long n_ffts, double* ms)
{
    if (n_ffts / 100 == 0) {
        if (ms) *ms = -1;
        return fail("smfft: FFT_multiple needs nFFTs >= 100");
    }
    Call c;
    if (r2c_call(&c, d_in, d_out, fft_size, n_ffts, 0, 100)) return 1;
    return timed(ms, c, t_stream);
}

// the repeated (FFT_multiple) code path with a caller-chosen repetition count, untimed: `reps` in-place transforms of each
// of the n_ffts transforms in d_in (NOT n_ffts/reps as the benchmark launchers count).  reps = 100 is the benchmark's
// instance; reps = 3 exists so that the path's VALUES can be checked (F^3 x stays finite).  mode 0 = C2C, 1 = R2C
// (forward; each repetition re-reads the packed spectrum as reals, as RC:374-376 does).
int smfft_exec_repeated(const void* d_in, void* d_out, int fft_size, long long n_ffts, int inverse, int reorder, int mode, int reps)
{
    if (reps != 3 && reps != 100) return fail("smfft: repeated instances exist for 3 and 100 repetitions");
    if (mode != 0 && mode != 1) return fail("smfft: repeated mode must be 0 (C2C) or 1 (R2C)");
    Call c;
    if (mode == 0 ? c2c_call(&c, d_in, d_out, fft_size, n_ffts * reps, inverse, reorder, reps) : r2c_call(&c, d_in, d_out, fft_size, n_ffts * reps, 0, reps)) return 1;
    return run_call(c, t_stream);
}

// ---- host-pointer drivers ---------------------------------------------------------------------

static int host_driver(int is_real, const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder,
                       int n_runs, double* single_ms, double* multi_ms)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    {   // validate before anything is allocated
        Call probe;
        if (is_real ? r2c_call(&probe, nullptr, nullptr, fft_size, n_ffts, inverse, 1) : c2c_call(&probe, nullptr, nullptr, fft_size, n_ffts, inverse, reorder, 1)) return 1;
    }
    if (n_ffts == 0) return 0;
    if (!h_in || !h_out) return fail("smfft: null host pointer");
    if (n_runs < 1) n_runs = 1;
    const cudaStream_t stream = t_stream;
    const size_t bytes = (size_t)n_ffts * (size_t)fft_size * (is_real ? sizeof(float) : sizeof(float2));
    size_t free_mem = 0, total_mem = 0;
    CUDA_TRY(cudaMemGetInfo(&free_mem, &total_mem));
    if (2 * bytes > free_mem) return fail_code(SMFFT_ERR_MEMORY, "smfft: not enough device memory (%zu B needed, %zu B free)", 2 * bytes, free_mem);  // CT:839-847
    void *d_in = nullptr, *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_in, bytes));
    if (cudaMalloc(&d_out, bytes) != cudaSuccess) { cudaFree(d_in); return fail_code(SMFFT_ERR_MEMORY, "smfft: cudaMalloc failed"); }
    int rc = 0;
    double t_multi = 0, t_single = 0;
    auto h2d = [&]() {
        cudaError_t e = cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, stream);
        return e == cudaSuccess ? 0 : fail_code(SMFFT_ERR_CUDA, "smfft: H2D copy failed: %s", cudaGetErrorString(e));
    };
    if (multi_ms && n_ffts >= 100 && !(is_real && inverse)) {
        Call c;
        rc = is_real ? r2c_call(&c, d_in, d_out, fft_size, n_ffts, 0, 100) : c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 100);
        for (int r = 0; r < n_runs && !rc; r++) {
            rc = h2d();
            if (!rc) rc = timed(&t_multi, c, stream);
        }
        *multi_ms = t_multi / n_runs;
    }
    Call c;
    if (!rc) rc = is_real ? r2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, 1) : c2c_call(&c, d_in, d_out, fft_size, n_ffts, inverse, reorder, 1);
    for (int r = 0; r < n_runs && !rc; r++) {
        rc = h2d();
        if (!rc) rc = timed(&t_single, c, stream);
    }
    if (single_ms) *single_ms = t_single / n_runs;
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) rc = fail_code(SMFFT_ERR_CUDA, "smfft: D2H copy failed: %s", cudaGetErrorString(e));
    }
    cudaError_t es = cudaStreamSynchronize(stream);
    if (!rc && es != cudaSuccess) rc = fail_code(SMFFT_ERR_CUDA, "smfft: stream sync failed: %s", cudaGetErrorString(es));
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
// events belong to the DEVICE (DeviceState::pipe): created on first use, kept between calls so repeated
// calls pay no allocation, released by smfft_pipeline_release().  Calls on one device serialise on its mutex.
namespace {
void pipeline_free_buffers(PipelineCtx& p)
{
    for (int i = 0; i < PipelineCtx::NBUF; i++) {
        if (p.d_in[i]) cudaFree(p.d_in[i]);
        if (p.d_out[i]) cudaFree(p.d_out[i]);
        p.d_in[i] = p.d_out[i] = nullptr;
    }
    p.cap = 0;
}

void pipeline_free_handles(PipelineCtx& p)
{
    if (p.s_in) cudaStreamDestroy(p.s_in);
    if (p.s_fft) cudaStreamDestroy(p.s_fft);
    if (p.s_out) cudaStreamDestroy(p.s_out);
    p.s_in = p.s_fft = p.s_out = nullptr;
    for (int i = 0; i < PipelineCtx::NBUF; i++) {
        if (p.ev_in[i]) cudaEventDestroy(p.ev_in[i]);
        if (p.ev_fft[i]) cudaEventDestroy(p.ev_fft[i]);
        if (p.ev_out[i]) cudaEventDestroy(p.ev_out[i]);
        p.ev_in[i] = p.ev_fft[i] = p.ev_out[i] = nullptr;
    }
    if (p.t0) cudaEventDestroy(p.t0);
    if (p.t1) cudaEventDestroy(p.t1);
    p.t0 = p.t1 = nullptr;
    p.handles = false;
}

// caller holds p.mu and has the owning device current
int pipeline_prepare(PipelineCtx& p, size_t chunk_bytes)
{
    if (!p.handles) {
        bool ok = cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&p.s_fft, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < PipelineCtx::NBUF && ok; i++)
            ok = cudaEventCreateWithFlags(&p.ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.ev_fft[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.ev_out[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreate(&p.t0) == cudaSuccess && cudaEventCreate(&p.t1) == cudaSuccess;
        if (!ok) {
            const cudaError_t e = cudaGetLastError();
            pipeline_free_handles(p);
            return fail_code(SMFFT_ERR_CUDA, "smfft: pipeline stream/event creation failed: %s", cudaGetErrorString(e));
        }
        p.handles = true;
    }
    if (p.cap >= chunk_bytes) return 0;
    pipeline_free_buffers(p);
    for (int i = 0; i < PipelineCtx::NBUF; i++) {
        if (cudaMalloc(&p.d_in[i], chunk_bytes) != cudaSuccess || cudaMalloc(&p.d_out[i], chunk_bytes) != cudaSuccess) {
            cudaGetLastError();
            pipeline_free_buffers(p);  // nothing half-allocated survives a failure
            return fail_code(SMFFT_ERR_MEMORY, "smfft: pipeline buffers (6 x %zu B) do not fit in device memory", chunk_bytes);
        }
    }
    p.cap = chunk_bytes;
    return 0;
}
}  // namespace

int smfft_pipeline_release(void)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    std::lock_guard<std::mutex> lk(ds->pipe.mu);
    if (ds->pipe.handles) {
        cudaStreamSynchronize(ds->pipe.s_in);
        cudaStreamSynchronize(ds->pipe.s_fft);
        cudaStreamSynchronize(ds->pipe.s_out);
    }
    pipeline_free_buffers(ds->pipe);
    pipeline_free_handles(ds->pipe);
    smfft::big::release();  // twiddle tables and scratch pool of the two-pass transforms
    return 0;
}

int smfft_pipeline_host(const void* h_in, void* h_out, int fft_size, long long n_ffts, int inverse, int reorder, int mode,
                        long long chunk_ffts, double* ms)
{
    DeviceState* ds = nullptr;
    if (get_device_state(&ds)) return 1;
    if (mode < 0 || mode > 2) return fail("smfft: pipeline mode must be 0 (C2C), 1 (R2C) or 2 (C2R)");
    {   // validate the size BEFORE it is used as a divisor
        Call probe;
        if (mode == 0 ? c2c_call(&probe, nullptr, nullptr, fft_size, n_ffts, inverse, reorder, 1) : r2c_call(&probe, nullptr, nullptr, fft_size, n_ffts, mode == 2, 1)) return 1;
    }
    if (n_ffts == 0) return 0;
    if (!h_in || !h_out) return fail("smfft: null host pointer");
    const size_t fft_bytes = (size_t)fft_size * (mode == 0 ? sizeof(float2) : sizeof(float));
    if (chunk_ffts <= 0) chunk_ffts = (long long)((size_t)g_pipe_chunk_bytes.load() / fft_bytes);
    if (chunk_ffts < 1) chunk_ffts = 1;
    if (chunk_ffts > n_ffts) chunk_ffts = n_ffts;
    PipelineCtx& p = ds->pipe;
    std::lock_guard<std::mutex> lk(p.mu);
    if (pipeline_prepare(p, (size_t)chunk_ffts * fft_bytes)) return 1;
    const int NBUF = PipelineCtx::NBUF;
    int rc = 0;
    cudaError_t ce = cudaSuccess;
#define PIPE_TRY(expr)                                                                                        \
    do {                                                                                                      \
        if (!rc && (ce = (expr)) != cudaSuccess) rc = fail_code(SMFFT_ERR_CUDA, "smfft: pipeline: %s: %s", #expr, cudaGetErrorString(ce)); \
    } while (0)
    PIPE_TRY(cudaEventRecord(p.t0, p.s_in));
    PIPE_TRY(cudaStreamWaitEvent(p.s_out, p.t0, 0));
    PIPE_TRY(cudaStreamWaitEvent(p.s_fft, p.t0, 0));
    long long done = 0;
    for (long long i = 0; done < n_ffts && !rc; i++, done += chunk_ffts) {
        const int b = (int)(i % NBUF);
        const long long cnt = (n_ffts - done) < chunk_ffts ? (n_ffts - done) : chunk_ffts;
        const char* src = (const char*)h_in + (size_t)done * fft_bytes;
        char* dst = (char*)h_out + (size_t)done * fft_bytes;
        if (i >= NBUF) {
            PIPE_TRY(cudaStreamWaitEvent(p.s_in, p.ev_fft[b], 0));   // d_in[b] consumed by the FFT of chunk i-NBUF
            PIPE_TRY(cudaStreamWaitEvent(p.s_fft, p.ev_out[b], 0));  // d_out[b] drained by the D2H of chunk i-NBUF
        }
        PIPE_TRY(cudaMemcpyAsync(p.d_in[b], src, cnt * fft_bytes, cudaMemcpyHostToDevice, p.s_in));
        PIPE_TRY(cudaEventRecord(p.ev_in[b], p.s_in));
        PIPE_TRY(cudaStreamWaitEvent(p.s_fft, p.ev_in[b], 0));
        if (!rc)
            rc = mode == 0 ? smfft_exec_c2c_stream(p.d_in[b], p.d_out[b], fft_size, cnt, inverse, reorder, p.s_fft)
                           : smfft_exec_r2c_c2r_stream(p.d_in[b], p.d_out[b], fft_size, cnt, mode == 2, p.s_fft);
        PIPE_TRY(cudaEventRecord(p.ev_fft[b], p.s_fft));
        PIPE_TRY(cudaStreamWaitEvent(p.s_out, p.ev_fft[b], 0));
        PIPE_TRY(cudaMemcpyAsync(dst, p.d_out[b], cnt * fft_bytes, cudaMemcpyDeviceToHost, p.s_out));
        PIPE_TRY(cudaEventRecord(p.ev_out[b], p.s_out));
    }
    PIPE_TRY(cudaEventRecord(p.t1, p.s_out));
#undef PIPE_TRY
    // drain all three streams whatever happened: the caller's buffers must not be touched after return
    const cudaError_t e1 = cudaStreamSynchronize(p.s_in), e2 = cudaStreamSynchronize(p.s_fft), e3 = cudaStreamSynchronize(p.s_out);
    const cudaError_t es = e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3;
    if (!rc && es != cudaSuccess) rc = fail_code(SMFFT_ERR_CUDA, "smfft: pipeline failed: %s", cudaGetErrorString(es));
    if (!rc && ms) {
        float t = 0;
        if (cudaEventElapsedTime(&t, p.t0, p.t1) == cudaSuccess) *ms += (double)t;
    }
    return rc;
}

}  // extern "C"
