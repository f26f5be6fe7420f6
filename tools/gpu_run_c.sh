#!/bin/bash
mkdir -p gpurun_out
echo "=== tune v2"; timeout 900 tools/tune 29 7 > gpurun_out/tune_c.csv 2> gpurun_out/tune_c.err; echo "tune rc=$?"; tail -2 gpurun_out/tune_c.err; wc -l gpurun_out/tune_c.csv
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_c.json 10 > gpurun_out/report_c.log 2>&1; echo "report rc=$?"; tail -2 gpurun_out/report_c.log | cut -c1-300
echo "=== sanitizer"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
