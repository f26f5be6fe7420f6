#!/bin/bash
# second on-box pass: parity with the measured shapes, full comparison report, ncu evidence, bench
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_b.log | tail -8
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_b.json 10 > gpurun_out/report_b.log 2>&1; echo "report rc=$?"; tail -4 gpurun_out/report_b.log | cut -c1-600
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ncu_launches_b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-baselines > gpurun_out/ncu_bench_b.log 2>&1; echo "ncu rc=$?"
echo "=== ncu full"
for cfg in "1024 1" "4096 1" "32 0"; do set -- $cfg
timeout 900 ncu --set full --clock-control none --import-source on -k regex:smfft_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_b_n$1_r$2 python tools/ncu_target.py $1 $2 > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu full $1 rc=$?"
done
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b.err; cut -c1-400 gpurun_out/bench_b.json
