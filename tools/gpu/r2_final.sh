#!/bin/bash
# final-state pass of the round: GPU tests, smoke, default bench line, reference arm, ncu launch list + metrics of the bench's
# launch list, ncu --set full of the dominant kernels, sanitizers
mkdir -p gpurun_out /tmp/ncu
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tee gpurun_out/r02_pytest_final.log | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_final.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['best_known']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['e2e']['frac'], d['clocks'], d['cpu_baseline']['value']); o=d['other_modes']; print({k:v['ms'] for k,v in o['r2c'].items()}); print({k:v['ms'] for k,v in o['c2r'].items()}); print({k:(v['ms'],v['TFLOPs']) for k,v in o['ct_multiple'].items()}); print(o['c2c_8192'], o['c2c_16384']); print(o['c2c_two_pass']); print(d['baselines']['steady_state_per_kernel']); print(d['device_api']['external']['worst_speedup'], d['device_api']['multiple']['worst_speedup'])"
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_ref.json 2>&1; tail -c 300 gpurun_out/r02_bench_final_ref.json
echo "=== ncu launch list (bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-baselines --no-other-modes --no-device-api > gpurun_out/r02_ncu_bench_final.log 2>&1; echo "rc=$?"
echo "=== ncu metrics"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:smfft_tile_kernel --csv --log-file gpurun_out/r02_ncu_metrics_final.csv python tools/ncu_metrics_target.py > gpurun_out/r02_ncu_metrics_final.log 2>&1; echo "ncu rc=$?"
echo "=== ncu full"
for cfg in "c2c 1024 1" "c2c 256 1" "c2c 4096 1" "c2c 16384 1" "c2c 32 1" "r2c 4096 1" "c2r 4096 1" "multiple 32 1" "multiple 1024 1"; do set -- $cfg
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smfft_tile_kernel -s 2 -c 1 -f -o /tmp/ncu/r02_final_$1_n$2_r$3 python tools/ncu_target.py $1 $2 $3 > gpurun_out/r02_ncu_full_final_$1_$2_$3.log 2>&1; echo "ncu full $1 $2 $3 rc=$?"
done
python tools/ncu_summarize.py gpurun_out/r02_ncu_summary_final.md /tmp/ncu/r02_final_*.ncu-rep > gpurun_out/r02_ncu_summarize_final.log 2>&1; echo "summarize rc=$?"
echo "=== sanitizer"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck_final.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_memcheck_final.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck_final.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_racecheck_final.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_synccheck_final.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_synccheck_final.log
du -sh gpurun_out
