#!/bin/bash
# final-state pass: GPU tests, bench line (incl. other_modes), launch list, full captures of the final kernels
# (summarised on the box: the .ncu-rep files of 9 captures exceed the 64 MiB that travels back), sanitizers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_zz.log | tail -4
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_zz.json 2> gpurun_out/bench_zz.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_zz.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['traffic']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['clocks'])"
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_zz_ref.json 2>&1; tail -c 300 gpurun_out/bench_zz_ref.json
echo "=== ncu launch list (bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ncu_launches_zz.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-baselines --no-other-modes > gpurun_out/ncu_bench_zz.log 2>&1; echo "rc=$?"
echo "=== ncu full"
mkdir -p /tmp/ncu
for cfg in "c2c 1024 1" "c2c 2048 1" "c2c 4096 1" "c2c 32 1" "c2c 128 1" "r2c 4096 1" "c2r 4096 1" "r2c 8192 1" "c2r 8192 1" "r2c 2048 1" "multiple 1024 1" "multiple 2048 1"; do set -- $cfg
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smfft_tile_kernel -s 2 -c 1 -f -o /tmp/ncu/prof_z_$1_n$2_r$3 python tools/ncu_target.py $1 $2 $3 > gpurun_out/ncu_full_zz_$1_$2_$3.log 2>&1; echo "ncu full $1 $2 $3 rc=$?"
done
python tools/ncu_summarize.py gpurun_out/ncu_summary_zz.md /tmp/ncu/prof_z_*.ncu-rep > gpurun_out/ncu_summarize_zz.log 2>&1; echo "summarize rc=$?"
ls -la /tmp/ncu; cp /tmp/ncu/prof_z_r2c_n4096_r1.ncu-rep /tmp/ncu/prof_z_c2c_n4096_r1.ncu-rep gpurun_out/ 2>/dev/null
echo "=== sanitizer"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck_zz.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_zz.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck_zz.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_zz.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py > gpurun_out/sanitizer_synccheck_zz.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/sanitizer_synccheck_zz.log
du -sh gpurun_out
