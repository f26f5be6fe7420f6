#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_j.log | tail -12
echo "=== tune"; timeout 1500 tools/tune 29 9 > gpurun_out/tune_j.csv 2> gpurun_out/tune_j.err; echo "rc=$?"; tail -2 gpurun_out/tune_j.err; wc -l gpurun_out/tune_j.csv
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_j.json 7 > gpurun_out/report_j.log 2>&1; echo "report rc=$?"; tail -2 gpurun_out/report_j.log | cut -c1-200
