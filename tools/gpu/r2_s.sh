#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest compat"; python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compat or native" 2>&1 | tail -2
echo "=== compat bench"; timeout 900 python tools/compat_bench.py gpurun_out/r02_compat_bench_final.json 7 > gpurun_out/r02_compat_bench_final.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02_compat_bench_final.log
echo "=== convolve bench"; timeout 600 python tools/convolve_bench.py gpurun_out/r02_convolve_bench_final.json 7 > gpurun_out/r02_convolve_bench_final.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_convolve_bench_final.log | cut -c1-400
