#!/bin/bash
# round 2: full GPU test suite on the current library, shuffle-exchange A/B for FFT_multiple, bench, ncu metrics, sanitizers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tee gpurun_out/r02_pytest_k.log | tail -5
echo "=== A/B shuffle exchange in FFT_multiple (A = shared-memory exchange, B = product)"; timeout 600 python tools/ab_multiple.py $PWD/smfft_b200/lib/libsmfft_noxshfl.so $PWD/smfft_b200/lib/libsmfft.so gpurun_out/r02_ab_xshfl_multiple.json 2>&1 | tail -7
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_k.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_k.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['best_known']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['clocks'], d['e2e']['value'], d['e2e']['frac']); print('cufft', {k:v['ms'] for k,v in d['baselines']['cufft_ms'].items() if isinstance(v,dict)}); print({k:v['ms'] for k,v in d['other_modes']['ct_multiple'].items()})"
echo "=== ncu metrics"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:smfft_tile_kernel --csv --log-file gpurun_out/r02_ncu_metrics_k.csv python tools/ncu_metrics_target.py > gpurun_out/r02_ncu_metrics_k.log 2>&1; echo "ncu rc=$?"
echo "=== sanitizer"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck_k.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_memcheck_k.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck_k.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_racecheck_k.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_synccheck_k.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_synccheck_k.log
du -sh gpurun_out
