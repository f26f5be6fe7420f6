#!/bin/bash
# round 2, second GPU pass: new GPU tests, compat engines (warp-shuffle) vs the reference kernels, cuFFT kernel shapes
mkdir -p gpurun_out /tmp/ncu
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tee gpurun_out/r02_pytest_b.log | tail -6
echo "=== compat bench"; timeout 900 python tools/compat_bench.py gpurun_out/r02_compat_bench_b.json 7 > gpurun_out/r02_compat_bench_b.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02_compat_bench_b.log
echo "=== cufft shapes"
timeout 600 ncu --metrics launch__grid_size,launch__block_size,launch__registers_per_thread,launch__shared_mem_per_block_static,launch__shared_mem_per_block_dynamic,launch__shared_mem_config_size,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fft --csv --log-file gpurun_out/r02_cufft_shapes.csv python tools/cufft_target.py > gpurun_out/r02_cufft_shapes.log 2>&1; echo "rc=$?"
echo "=== ncu full: compat vs reference, N=1024"
for cfg in "compat external 1024 0 1" "compat external 1024 0 0" "reference external 1024 0 1" "reference external 1024 0 0" "compat multiple 1024 0 1" "compat multiple 1024 0 0" "reference multiple 1024 0 1" "compat multiple 32 0 1" "reference multiple 32 0 1"; do set -- $cfg
timeout 600 ncu --set full --clock-control none --import-source on -k regex:SMFFT_DIT -s 2 -c 1 -f -o /tmp/ncu/r02_$1_$2_n$3_i$4_r$5 python tools/compat_target.py $1 $2 $3 $4 $5 > gpurun_out/r02_ncu_$1_$2_$3_$4_$5.log 2>&1; echo "ncu $cfg rc=$?"
done
python tools/ncu_summarize.py gpurun_out/r02_ncu_summary_compat_b.md /tmp/ncu/r02_*.ncu-rep > gpurun_out/r02_ncu_summarize_b.log 2>&1; echo "summarize rc=$?"
du -sh gpurun_out
