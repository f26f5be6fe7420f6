#!/bin/bash
# evidence pass: ncu launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out
echo "=== ncu launch list (bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/ncu_launches_s.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-baselines > gpurun_out/ncu_bench_s.log 2>&1; echo "rc=$?"
echo "=== ncu full"
for cfg in "c2c 1024 1" "c2c 4096 1" "c2c 32 0" "c2c 256 1" "r2c 2048 1" "c2r 2048 1" "multiple 1024 1" "multiple 32 1"; do set -- $cfg
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smfft_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_s_$1_n$2_r$3 python tools/ncu_target.py $1 $2 $3 > gpurun_out/ncu_full_s_$1_$2.log 2>&1; echo "ncu full $1 $2 rc=$?"
done
