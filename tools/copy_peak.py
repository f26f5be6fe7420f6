#!/usr/bin/env python
"""tools/copy_peak.py -- the host<->device copy ceiling the end-to-end path (smfft_pipeline_host) runs against.

Bare cudaMemcpyAsync on pinned host buffers, no FFT: H2D alone, D2H alone, and both directions concurrently on two
streams (what the pipeline does), whole-buffer and in 128 MiB chunks; optionally write-combined pinned memory for the
H2D source.  Under torchrun every rank measures at the same time (barrier before each case), so the per-rank and the
aggregate host limits at 1/2/4/8 ranks come out of the same tool.  GB = 1e9 bytes.

    python tools/copy_peak.py out.json [GiB]
    python -m torch.distributed.run --nproc-per-node N ... tools/copy_peak.py out.json
"""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/copy_peak.json"
    gib = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = int(gib * (1 << 30))
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(0)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cudart = ctypes.CDLL("libcudart.so.12")
    cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def h2d(chunk, src=None):
        src = src if src is not None else h_in.data_ptr()
        for o in range(0, nbytes, chunk):
            cudart.cudaMemcpyAsync(d_in.data_ptr() + o, src + o, min(chunk, nbytes - o), 1, s1.cuda_stream)

    def d2h(chunk):
        for o in range(0, nbytes, chunk):
            cudart.cudaMemcpyAsync(h_out.data_ptr() + o, d_out.data_ptr() + o, min(chunk, nbytes - o), 2, s2.cuda_stream)

    def timed(fns, reps=3):
        best = None
        for _ in range(reps + 1):
            barrier()
            t0 = time.perf_counter()
            for f in fns:
                f()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = dt if best is None else min(best, dt)
        return best

    res = {"world": world, "bytes_per_direction_per_rank": nbytes, "cases": {}}
    for name, chunk in (("whole", nbytes), ("chunk128MiB", 128 << 20), ("chunk32MiB", 32 << 20)):
        t_h = timed([lambda: h2d(chunk)])
        t_d = timed([lambda: d2h(chunk)])
        t_b = timed([lambda: h2d(chunk), lambda: d2h(chunk)])
        res["cases"][name] = {"h2d_only_GBps_per_rank": nbytes / t_h / 1e9, "d2h_only_GBps_per_rank": nbytes / t_d / 1e9,
                              "both_GBps_per_direction_per_rank": nbytes / t_b / 1e9,
                              "both_GBps_aggregate_in_plus_out": 2 * nbytes * world / t_b / 1e9}
    # write-combined pinned source for H2D (cudaHostAllocWriteCombined = 4)
    try:
        p = ctypes.c_void_p()
        cudart.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
        if cudart.cudaHostAlloc(ctypes.byref(p), nbytes, 4) == 0:
            ctypes.memset(p, 1, nbytes)
            t_h = timed([lambda: h2d(128 << 20, p.value)])
            t_b = timed([lambda: h2d(128 << 20, p.value), lambda: d2h(128 << 20)])
            res["cases"]["chunk128MiB_wc_source"] = {"h2d_only_GBps_per_rank": nbytes / t_h / 1e9,
                                                     "both_GBps_per_direction_per_rank": nbytes / t_b / 1e9,
                                                     "both_GBps_aggregate_in_plus_out": 2 * nbytes * world / t_b / 1e9}
            cudart.cudaFreeHost(p)
    except Exception as ex:  # pragma: no cover
        res["cases"]["wc_error"] = str(ex)[:200]
    try:
        res["affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]
    except Exception:
        pass
    if rank == 0:
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
