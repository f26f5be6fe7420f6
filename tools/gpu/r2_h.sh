#!/bin/bash
# per-size DRAM traffic and pipe utilisation of the final kernels (ncu metrics pass, full-size batch) + multi-arm bench
mkdir -p gpurun_out
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:smfft_tile_kernel --csv --log-file gpurun_out/r02_ncu_metrics_h.csv python tools/ncu_metrics_target.py > gpurun_out/r02_ncu_metrics_h.log 2>&1; echo "ncu rc=$?"
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_h.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_h.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['best_known']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['clocks']); print('cufft', {k:v['ms'] for k,v in d['baselines']['cufft_ms'].items() if isinstance(v,dict)})"
