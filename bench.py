#!/usr/bin/env python
"""bench.py -- headline benchmark of the SMFFT hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a kernels
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: oracle port on the host cores

Workload (BASELINE.json configs[1]): Cooley-Tukey C2C, N = 32..4096, reorder and no-reorder, forward,
one 4 GiB float2 batch per GPU (16M x 32 ... 131k x 4096).  One "step" = the 16 FFT_external launches
(8 sizes x 2 reorder modes) over that batch.  Metric: HBM GB/s = algorithmic bytes (16 B per complex
point) / time; whole-job value = bytes of all ranks / max-over-ranks time.  Independent
FFTs shard by batch: every rank transforms its own 4 GiB (weak scaling), no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096]
BATCH_POINTS = 1 << 29          # 4 GiB of float2
BYTES_PER_POINT = 16            # 8 read + 8 written
METRIC = "Batched FFT HBM GB/s & ms per 4 GB batch, N=32\u20134096, at 1/2/4/8 B200"  # BASELINE.json metric, verbatim


def configs():
    return [(n, reorder) for n in SIZES for reorder in (1, 0)]


def reduce_job(t_ms, units):
    """job time = max over ranks, job units = sum over ranks (torch.distributed; works on gloo and nccl)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    return float(t_ms.item()), float(units.item())


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    @staticmethod
    def _epoch(stamp):
        import datetime

        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t_begin=None, t_end=None):
        """the sampler is started early (nvidia-smi needs a few hundred ms to come up on an 8-GPU box); samples are kept
        when their timestamp falls inside [t_begin, t_end] (the timed region), all of them if that leaves nothing"""
        if self.p is not None:
            time.sleep(0.06)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        text = self.f.read()
        if not text.strip():  # the looping sampler produced nothing (very short region / busy nvidia-smi): one direct sample
            try:
                text = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                      capture_output=True, text=True, timeout=20).stdout
            except Exception:
                text = ""
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in text.splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                rows.append((self._epoch(c[9]) if len(c) > 9 else None, float(c[1]), float(c[2]),
                             [name for name, v in zip(names, c[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        if t_begin is not None and t_end is not None:
            inside = [r for r in rows if r[0] is not None and t_begin - 0.03 <= r[0] <= t_end + 0.03]
            if inside:
                rows = inside
        sm, mx, reasons = [r[1] for r in rows], [r[2] for r in rows], set(x for r in rows for x in r[3])
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement (the reference has no CPU transform)
# ----------------------------------------------------------------------------------------------
def cpu_sweep(sample_points: int, repeats: int = 1):
    """times the oracle port (all host threads) on `sample_points` complex points per configuration.
    Returns (GB/s, seconds, cores)."""
    import numpy as np

    from oracle import oracle_np as O

    lib = O.c_oracle()
    lib.oracle_set_num_threads(os.cpu_count() or 1)  # all host threads (torchrun exports OMP_NUM_THREADS=1)
    cores = lib.oracle_num_threads()
    x = O.uniform_c64(1, sample_points).reshape(-1)
    out = np.empty_like(x)
    t0 = time.perf_counter()
    for _ in range(repeats):
        for n, reorder in configs():
            lib.oracle_ct_c2c_f32(x.ctypes.data, out.ctypes.data, n, sample_points // n, 0, reorder, 0)
    dt = time.perf_counter() - t0
    nbytes = repeats * len(configs()) * sample_points * BYTES_PER_POINT
    return nbytes / dt / 1e9, dt, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    sample = 1 << 20
    for _ in range(args.warmup):
        cpu_sweep(sample)
    t0 = time.perf_counter()
    vals = [cpu_sweep(sample) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    gbs = args.steps * len(configs()) * sample * BYTES_PER_POINT / dt / 1e9
    cores = vals[0][2]
    sample_txt = f"{len(configs())} configs x {sample} complex points per step (8 MiB in + 8 MiB out each), fp32 radix-2 DIT restatement"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CT C2C forward N=32..4096 x {reorder,no-reorder}; CPU arm on a bounded sample of the 4 GiB batch",
                   "note": "the reference ships no CPU transform (FFT.c only checks against cuFFT); this arm is the oracle's C restatement with OpenMP"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# reported GPU baselines (outside the timed region, rank 0): recompiled reference kernels, cuFFT
# ----------------------------------------------------------------------------------------------
def sustained_arm(launchers, steps, warmup=3):
    """The protocol of the timed region, for any arm: `steps` back-to-back passes over the whole launcher list (the
    16-launch step), every launch between two CUDA events on the launching stream; per key the median / min over the
    steps.  launchers: [(key, callable)]; a key may repeat inside a step (cuFFT has no no-reorder mode)."""
    import torch

    for _ in range(warmup):
        for _, fn in launchers:
            fn()
    torch.cuda.synchronize()
    rec = {}
    for _ in range(steps):
        for key, fn in launchers:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            rec.setdefault(key, []).append((e0, e1))
    torch.cuda.synchronize()
    out = {}
    for key, evs in rec.items():
        ts = [a.elapsed_time(b) for a, b in evs]
        out[key] = {"ms": round(statistics.median(ts), 4), "ms_min": round(min(ts), 4)}
    return out


def reference_gpu_baseline(x, y, steps=5, warmup=3):
    """the reference's own kernels rebuilt for sm_100a (oracle/_ref), through ITS FFT_external_benchmark (CT:583), in the
    same sustained 16-launch step as the product"""
    import ctypes

    so = os.path.join(ROOT, "oracle", "_ref", "libsmfft_ref_ct.so")
    if not os.path.exists(so):
        return None
    try:
        ref = ctypes.CDLL(so)
        fn = getattr(ref, "_Z22FFT_external_benchmarkP6float2S0_iibbPd")
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_bool, ctypes.c_bool,
                       ctypes.POINTER(ctypes.c_double)]
        ms = ctypes.c_double(0)
        xp, yp = x.data_ptr(), y.data_ptr()
        launchers = [(f"{n}{'r' if reorder else 'n'}", (lambda n=n, reorder=reorder: fn(xp, yp, n, BATCH_POINTS // n, False, bool(reorder), ctypes.byref(ms))))
                     for n, reorder in configs()]
        out = sustained_arm(launchers, steps, warmup)
        out["how"] = "unmodified reference kernels (sm_100a rebuild), sustained 16-launch step, CUDA events per launch, median over %d steps after %d warm-up steps" % (steps, warmup)
        return out
    except Exception as ex:  # pragma: no cover
        return {"error": str(ex)[:200]}


def cufft_baseline(x, y, steps=5, warmup=3):
    """cuFFT on the same buffers: plans created outside the timer, out of place, one cufftExecC2C per timed
    launch, in the same sustained 16-launch step as the product (cuFFT has no no-reorder mode: each size runs twice per step)."""
    import ctypes

    lib = None
    for name in ("libcufft.so.11", "libcufft.so.12", "libcufft.so"):
        try:
            lib = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if lib is None:
        return {"error": "libcufft not loadable"}
    plans = {}
    try:
        for n in SIZES:
            h = ctypes.c_int(0)
            if lib.cufftPlan1d(ctypes.byref(h), n, 0x29, BATCH_POINTS // n) == 0:
                plans[n] = h
        xp, yp = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr())
        launchers = [(str(n), (lambda n=n: lib.cufftExecC2C(plans[n], xp, yp, -1))) for n in SIZES if n in plans for _ in (0, 1)]
        out = sustained_arm(launchers, steps, warmup)
        out["how"] = ("cufftPlan1d(C2C, batch) + cufftExecC2C, out of place, sustained 16-launch step (each size twice), CUDA events per launch, "
                      "median over %d steps after %d warm-up steps -- the SAME step count and warm-up as the timed region (a burst shorter than ~50 ms runs "
                      "at higher SM clocks than the steady state the headline is measured in, profiles/r02_burst_timeline_*.json)" % (steps, warmup))
        return out
    except Exception as ex:  # pragma: no cover
        return {"error": str(ex)[:200]}
    finally:
        for h in plans.values():
            lib.cufftDestroy(h)


def steady_state_per_kernel(x, y, sizes=(128, 256, 512, 1024), count=100, tail=20):
    """ONE kernel at a time, `count` back-to-back launches from an idle GPU, mean of the last `tail` launches (per-launch CUDA
    events): each kernel in its OWN thermal steady state, free of what its neighbours in a mixed step draw.  Inside the mixed
    16-launch step a kernel inherits the clocks its neighbours leave behind: cuFFT's 32 / 64 / 2048 / 4096-point kernels run at
    4.3-5.8 TB/s and keep its 512 / 1024-point kernels cooler than ours get to be (profiles/r02_burst_timeline_p.json)."""
    import ctypes

    import torch

    import smfft_b200 as sm

    out = {"how": f"{count} back-to-back launches of one kernel after 0.5 s idle, mean of the last {tail}, CUDA events per launch; natural-order C2C, 4 GiB batch"}
    try:
        cu = ctypes.CDLL("libcufft.so.11")
    except OSError:
        cu = None
    xp, yp = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr())
    for n in sizes:
        arms = {"ours": lambda: sm.exec_c2c(x, y, n, BATCH_POINTS // n, False, True)}
        h = ctypes.c_int(0)
        if cu is not None and cu.cufftPlan1d(ctypes.byref(h), n, 0x29, BATCH_POINTS // n) == 0:
            arms["cufft"] = lambda: cu.cufftExecC2C(h, xp, yp, -1)
        row = {}
        for name, fn in arms.items():
            fn()
            torch.cuda.synchronize()
            time.sleep(0.5)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(count + 1)]
            ev[0].record()
            for i in range(count):
                fn()
                ev[i + 1].record()
            torch.cuda.synchronize()
            ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(count)]
            row[name] = {"first5_ms": round(sum(ts[:5]) / 5, 4), "steady_ms": round(sum(ts[-tail:]) / tail, 4)}
        if "cufft" in row:
            row["ours_vs_cufft_steady"] = round(row["cufft"]["steady_ms"] / row["ours"]["steady_ms"], 4)
            cu.cufftDestroy(h)
        out[str(n)] = row
    return out


def device_api_leg(x, y, rounds=5):
    """The reference's DEVICE API (the product surface a user kernel calls, README.md:10-20 of the reference) on this
    library vs on the reference itself, same launch shapes, same buffers, launched alternately: SMFFT_DIT_external<P> and
    SMFFT_DIT_multiple<P> from include/smfft/compat.cuh (tests/compat/compat_kernels.cu) against the same-named kernels of the
    unmodified reference (oracle/_ref).  The `multiple` ratio is the cost of the in-shared-memory transform a user kernel
    pays (100 in-place calls of do_SMFFT_CT_DIT per tile).  Full table: tools/compat_bench.py -> profiles/."""
    import ctypes

    try:
        from tests.compat.build_compat import build as build_compat

        lib = ctypes.CDLL(build_compat())
        ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsmfft_ref_ct.so"))
    except Exception as ex:  # pragma: no cover
        return {"error": str(ex)[:200]}
    P, I, B, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_bool, ctypes.POINTER(ctypes.c_double)
    for name in ("compat_ct_external", "compat_ct_multiple"):
        getattr(lib, name).argtypes = [P, P, I, I, I, I]
    rext = getattr(ref, "_Z22FFT_external_benchmarkP6float2S0_iibbPd")
    rmul = getattr(ref, "_Z22FFT_multiple_benchmarkP6float2S0_iibbPd")
    rext.argtypes = rmul.argtypes = [P, P, I, I, B, B, D]
    import torch

    ms = ctypes.c_double(0)
    xp, yp = x.data_ptr(), y.data_ptr()
    out = {}

    def once(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    try:
        for kind, cfn, rfn in (("external", lib.compat_ct_external, rext), ("multiple", lib.compat_ct_multiple, rmul)):
            rows = {}
            for n, r in configs():
                fa = lambda: cfn(xp, yp, n, BATCH_POINTS // n, 0, r)
                fb = lambda: rfn(xp, yp, n, BATCH_POINTS // n, False, bool(r), ctypes.byref(ms))
                fa(), fb()
                ta, tb = [], []
                for _ in range(rounds):   # interleaved: both kernels see the same clocks (inside separate arms each inherits its neighbours')
                    ta.append(once(fa))
                    tb.append(once(fb))
                a, b = statistics.median(ta), statistics.median(tb)
                rows[f"{n}{'r' if r else 'n'}"] = {"compat_ms": round(a, 4), "reference_ms": round(b, 4), "speedup": round(b / a, 3)}
            sp = [v["speedup"] for v in rows.values()]
            out[kind] = {"per_size": rows, "worst_speedup": min(sp), "median_speedup": round(statistics.median(sp), 3)}
        out["how"] = ("SMFFT_DIT_external / SMFFT_DIT_multiple<FFT_N_forward[_noreorder]>, reference launch shapes, 4 GiB batch, compat and reference "
                      "launched alternately (A B A B ...), CUDA events per launch, median of %d; speedup = reference ms / compat ms; the one-warp tiles "
                      "(N <= 128, external) are bound by the CTA launch rate of the reference's own launch shape for both (DESIGN.md 9.1)" % rounds)
    except Exception as ex:  # pragma: no cover
        out["error"] = str(ex)[:200]
    return out


def other_modes(x, y, reps=5):
    """The other BASELINE.json configurations on the same 4 GiB buffers, outside the timed region (reported only):
    configs[2] Stockham C2C forward+inverse N = 256..4096, configs[3] R2C / C2R on 4 GiB of reals (real N = 64..8192),
    configs[4] FFT_multiple (100 in-place repetitions per tile, compute-bound; flops = 5 N log2 N per transform).
    ms = median of `reps` launches timed by the library's own CUDA events; GB/s from the algorithmic bytes
    (16 B per complex point, 8 B per real point)."""
    import math

    import torch

    import smfft_b200 as sm

    peak, _ = measured_peak()

    def med(fn):
        fn()
        return statistics.median([fn() for _ in range(reps)])

    def row(ms, nbytes):
        gbs = nbytes / ms / 1e6
        return {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}

    out = {"stockham_c2c": {}, "r2c": {}, "c2r": {}, "ct_multiple": {}}
    try:
        for n in (256, 512, 1024, 2048, 4096):
            nf = BATCH_POINTS // n
            out["stockham_c2c"][str(n)] = {
                "forward": row(med(lambda: sm.Stockham_external_benchmark(x, y, n, nf, False)), BATCH_POINTS * 16),
                "inverse": row(med(lambda: sm.Stockham_external_benchmark(x, y, n, nf, True)), BATCH_POINTS * 16)}
        for nbig in (8192, 16384):   # beyond the reference: one transform per 64 / 128 KB tile
            out[f"c2c_{nbig}"] = {("reorder" if r else "noreorder"): row(med(lambda: sm.FFT_external_benchmark(x, y, nbig, BATCH_POINTS // nbig, False, bool(r))), BATCH_POINTS * 16)
                                  for r in (1, 0)}
        # 2^15 .. 2^24 points: two or three passes over HBM (csrc/big_fft.cu) -- 32 / 48 algorithmic bytes per point; cuFFT's
        # plan for the same batch on the same buffers beside it
        out["c2c_two_pass"] = {"how": "FFT_external_benchmark, natural order, 4 GiB batch, scratch chunk 1 GiB; two passes up to 2^20 points, three from 2^21; "
                                      "frac = 32 (48) B/point over the measured copy peak"}
        try:
            import ctypes
            cu = ctypes.CDLL("libcufft.so.11")
        except OSError:
            cu = None
        for nbig in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 20, 1 << 24):
            ms = med(lambda: sm.FFT_external_benchmark(x, y, nbig, BATCH_POINTS // nbig, False, True))
            bpp = 32 if nbig <= (1 << 20) else 48
            r = {"ms": round(ms, 4), "passes": bpp // 16, "GBps_traffic": round(BATCH_POINTS * bpp / ms / 1e6, 1), "frac": round(BATCH_POINTS * bpp / ms / 1e6 / peak, 4)}
            if cu is not None:
                h = ctypes.c_int(0)
                if cu.cufftPlan1d(ctypes.byref(h), nbig, 0x29, BATCH_POINTS // nbig) == 0:
                    xp, yp = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr())

                    def cufft_once():
                        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                        ev[0].record()
                        cu.cufftExecC2C(h, xp, yp, -1)
                        ev[1].record()
                        torch.cuda.synchronize()
                        return ev[0].elapsed_time(ev[1])

                    r["cufft_ms"] = round(med(cufft_once), 4)
                    r["ours_vs_cufft"] = round(r["cufft_ms"] / ms, 4)
                    cu.cufftDestroy(h)
            out["c2c_two_pass"][str(nbig)] = r
        real_points = 2 * BATCH_POINTS          # the same 4 GiB read as floats
        for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
            nf = real_points // n
            out["r2c"][str(n)] = row(med(lambda: sm.R2C_C2R_external_benchmark(x, y, n, nf, 0)), real_points * 8)
            out["c2r"][str(n)] = row(med(lambda: sm.R2C_C2R_external_benchmark(x, y, n, nf, 1)), real_points * 8)
        for n in SIZES:
            nf = BATCH_POINTS // n              # the launcher transforms nFFTs/100 tiles' worth of data 100 times (CT:669)
            for reorder in (1, 0):
                ms = med(lambda: sm.FFT_multiple_benchmark(x, y, n, nf, False, bool(reorder)))
                flops = 5.0 * n * math.log2(n) * (nf // 100) * 100
                out["ct_multiple"][f"{n}{'r' if reorder else 'n'}"] = {"ms": round(ms, 4), "TFLOPs": round(flops / ms / 1e9, 2)}
    except Exception as ex:  # pragma: no cover
        out["error"] = str(ex)[:200]
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import smfft_b200 as sm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: everything native libraries print there (NCCL's version banner, the
    # reference kernels' own printf) is sent to stderr at the file-descriptor level; the line goes out on the saved fd
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sm.FFT_init()

    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260101 + rank)
    x = torch.rand((BATCH_POINTS, 2), dtype=torch.float32, device="cuda", generator=gen)  # uniform [0,1) like CT/FFT.c:139-143
    y = torch.empty_like(x)
    cfgs = configs()

    def step(record=None):
        for n, reorder in cfgs:
            if record is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            sm.exec_c2c(x, y, n, BATCH_POINTS // n, False, bool(reorder))
            if record is not None:
                e1.record()
                record.append(((n, reorder), e0, e1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = sm.launch_count()
    rec = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall_begin = time.time()
    t0.record()
    for _ in range(args.steps):
        step(rec)
    t1.record()
    barrier()
    wall_end = time.time()
    launches = sm.launch_count() - launches0
    clocks = sampler.stop(wall_begin, wall_end)
    ms_total = t0.elapsed_time(t1)
    step_bytes = len(cfgs) * BATCH_POINTS * BYTES_PER_POINT
    tmax, units = reduce_job(torch.tensor([ms_total], dtype=torch.float64, device="cuda"),
                             torch.tensor([float(step_bytes * args.steps)], dtype=torch.float64, device="cuda"))
    value = units / (tmax * 1e-3) / 1e9

    per = {}
    for key, e0, e1 in rec:
        per.setdefault(key, []).append(e0.elapsed_time(e1))
    per_size = {f"{n}{'r' if r else 'n'}": {"ms": round(statistics.median(v), 4), "ms_min": round(min(v), 4),
                                             "GBps": round(BATCH_POINTS * BYTES_PER_POINT / statistics.median(v) / 1e6, 1)}
                for (n, r), v in per.items()}
    all_ms = [t for v in per.values() for t in v]
    avg_launch_ms = sum(all_ms) / len(all_ms)
    peak, peak_src = measured_peak()
    achieved = BATCH_POINTS * BYTES_PER_POINT / (avg_launch_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the final kernels: ncu metrics pass over this very launch list on the 4 GiB batch
    # (tools/ncu_metrics_target.py -> profiles/r02_ncu_metrics_*.csv -> tools/ncu_metrics_parse.py -> this file), per size
    traffic, traffic_detail, ncu_multiple = None, None, {}
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_detail = {"source": tj.get("source"), "per_size_over_algorithmic": {k: round(v["dram_bytes"] / (BATCH_POINTS * BYTES_PER_POINT), 4)
                                                                                      for k, v in tj.get("per_size", {}).items()}}
            ncu_multiple = tj.get("multiple", {})
        except Exception:
            traffic = None

    # ---- e2e: host buffers, H2D and D2H inside the timed region, through the C ABI pipeline ----
    e2e = None
    if not args.no_e2e:
        try:
            hx = torch.empty((BATCH_POINTS, 2), dtype=torch.float32, pin_memory=True)
            hy = torch.empty((BATCH_POINTS, 2), dtype=torch.float32, pin_memory=True)
        except RuntimeError:  # host cannot pin 8 GiB per rank: pageable buffers (slower copies, same path)
            hx = torch.empty((BATCH_POINTS, 2), dtype=torch.float32)
            hy = torch.empty((BATCH_POINTS, 2), dtype=torch.float32)
        hx.copy_(x)  # same synthetic batch, now host-resident
        torch.cuda.synchronize()
        # the ceiling e2e runs against: bare copies of the same pinned buffers, H2D and D2H concurrently on two streams
        # (what the pipeline overlaps), no FFT; every rank at the same time, max over ranks -- the host-side limit at N ranks
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        xf, yf, hxf, hyf = x.view(-1), y.view(-1), hx.view(-1), hy.view(-1)

        def bare_copy(chunk_elems):
            barrier()
            c0 = time.perf_counter()
            for o in range(0, xf.numel(), chunk_elems):
                with torch.cuda.stream(s_h2d):
                    yf[o:o + chunk_elems].copy_(hxf[o:o + chunk_elems], non_blocking=True)
                with torch.cuda.stream(s_d2h):
                    hyf[o:o + chunk_elems].copy_(xf[o:o + chunk_elems], non_blocking=True)
            torch.cuda.synchronize()
            return (time.perf_counter() - c0) * 1e3

        copy_ms = None
        for chunk in (xf.numel(), (128 << 20) // 4, (32 << 20) // 4):   # whole buffer, 128 MiB and 32 MiB pieces: the ceiling is the best any of them does
            for _ in range(3):       # the first pass touches the pinned pages
                t_ms = bare_copy(chunk)
                copy_ms = t_ms if copy_ms is None else min(copy_ms, t_ms)
        cmax, cunits = reduce_job(torch.tensor([copy_ms], dtype=torch.float64, device="cuda"),
                                  torch.tensor([float(2 * BATCH_POINTS * 8)], dtype=torch.float64, device="cuda"))
        copy_peak = cunits / (cmax * 1e-3) / 1e9
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        sm.pipeline_host(hx, hy, 1024, BATCH_POINTS // 1024, False, True)  # warm-up (allocations, pinned pages)
        barrier()
        w0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(e2e_steps):
            for n, reorder in cfgs:
                dev_ms += sm.pipeline_host(hx, hy, n, BATCH_POINTS // n, False, bool(reorder))
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - w0) * 1e3
        tmax2, units2 = reduce_job(torch.tensor([wall_ms], dtype=torch.float64, device="cuda"),
                                   torch.tensor([float(step_bytes * e2e_steps)], dtype=torch.float64, device="cuda"))
        e2e = {"value": units2 / (tmax2 * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": len(cfgs) * BATCH_POINTS * 8,
               "d2h_bytes_per_step": len(cfgs) * BATCH_POINTS * 8, "steps": e2e_steps, "ms_per_step": tmax2 / e2e_steps,
               "device_event_ms_per_step": dev_ms / e2e_steps,
               "api": "smfft_pipeline_host (pinned host buffers, chunked H2D->FFT->D2H on 3 streams), wall clock incl. allocation",
               "peak": copy_peak, "frac": (units2 / (tmax2 * 1e-3) / 1e9) / copy_peak,
               "peak_how": "bare cudaMemcpyAsync of the same pinned 4 GiB buffers, H2D and D2H concurrently on two streams (whole buffer, 128 MiB and "
                           "32 MiB pieces; best of 9), all ranks at once, bytes in + bytes out over the max-over-ranks wall time (GB/s, same unit as "
                           "value); both numbers are bound by the same PCIe / host path, so frac sits near 1 with a few percent of run-to-run spread"}
        sm.pipeline_release()  # the pipeline's device buffers (6 x 128 MiB) are not needed by the legs that follow
        del hx, hy

    cpu = None
    baselines = {}
    others = None
    device_api = None
    if rank == 0:
        if not args.no_other_modes:
            others = other_modes(x, y)
            # achieved FP32-pipe utilisation of FFT_multiple (north_star): from the same ncu metrics pass, per configuration
            for k, v in ncu_multiple.items():
                if isinstance(others.get("ct_multiple"), dict) and k in others["ct_multiple"]:
                    others["ct_multiple"][k]["fma_pipe_pct_ncu"] = v.get("fma_pipe_pct")
                    others["ct_multiple"][k]["smem_pipe_pct_ncu"] = v.get("smem_pipe_pct")
        if world == 1 and not args.no_cpu:
            sample = 1 << 20
            cpu_sweep(sample)
            reps = 1
            gbs, dt, cores = cpu_sweep(sample, reps)
            while dt < 10.0 and reps < 64:
                reps *= 2
                gbs, dt, cores = cpu_sweep(sample, reps)
            cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                   "sample": f"{reps} x {len(cfgs)} configs x {sample} complex points ({dt:.1f} s), oracle C restatement of CT:334-532 with OpenMP"}
        if not args.no_baselines:
            baselines["cufft_ms"] = cufft_baseline(x, y, args.steps, max(args.warmup, 3))
            baselines["reference_sm100a_ms"] = reference_gpu_baseline(x, y, max(3, args.steps // 2), 2)
            try:
                baselines["steady_state_per_kernel"] = steady_state_per_kernel(x, y)
            except Exception as ex:  # pragma: no cover
                baselines["steady_state_per_kernel"] = {"error": str(ex)[:200]}
        if not args.no_device_api:
            device_api = device_api_leg(x, y)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    # the best rate anything has demonstrated on these buffers under this protocol (cuFFT's fastest size), and ours against it
    best_known = None
    cu = baselines.get("cufft_ms") if isinstance(baselines, dict) else None
    if isinstance(cu, dict):
        cms = {k: v["ms"] for k, v in cu.items() if isinstance(v, dict) and "ms" in v}
        if cms:
            kbest = min(cms, key=cms.get)
            bk = BATCH_POINTS * BYTES_PER_POINT / cms[kbest] / 1e6
            ours_same = per_size.get(f"{kbest}r", {}).get("ms")
            best_known = {"GBps": round(bk, 1), "what": f"cuFFT C2C N={kbest}, same buffers, same sustained step", "frac_of_it": round(achieved / bk, 4),
                          "ours_same_size_ms": ours_same, "cufft_ms": cms[kbest],
                          "ours_vs_cufft_per_size": {k: round(cms[k] / per_size[f"{k}r"]["ms"], 4) for k in cms if f"{k}r" in per_size}}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": tmax / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CT C2C forward, N=32..4096 x {reorder, no-reorder}: 16 FFT_external launches per step, 4 GiB float2 batch per GPU (BASELINE.json configs[1])",
                   "batch_bytes_in": BATCH_POINTS * 8, "l2": "inputs (4 GiB) larger than L2 (126 MB); no flush needed",
                   "sharding": f"batch-sharded x{world}, no data-path collective",
                   "io": {0: "auto (measured best staging per size)", 1: "ldg", 2: "tma", 3: "tma_stg", 4: "reg"}[sm.get_option("io")], "twiddle": "lut" if sm.get_option("twiddle") == 0 else "mufu"},
        "ms_per_4GiB_batch": avg_launch_ms,
        "per_size": per_size,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "smfft_tile_kernel (mean over the 16 instances of a step)",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": BATCH_POINTS * BYTES_PER_POINT,
                     "traffic_detail": traffic_detail,
                     "frac_of_nominal_8TBps": achieved / 8000.0, "frac_of_hgx_7p7TBps": achieved / 7700.0,
                     "best_known": best_known},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "other_modes": others, "baselines": baselines, "device_api": device_api,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-baselines", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true")
    ap.add_argument("--no-device-api", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
