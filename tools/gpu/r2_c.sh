#!/bin/bash
# round 2, third GPU pass: native primitive tests, compat engines with packed butterflies, fused convolution four ways, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tee gpurun_out/r02_pytest_c.log | tail -6
echo "=== compat bench"; timeout 900 python tools/compat_bench.py gpurun_out/r02_compat_bench_c.json 7 > gpurun_out/r02_compat_bench_c.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02_compat_bench_c.log
echo "=== convolve bench"; timeout 600 python tools/convolve_bench.py gpurun_out/r02_convolve_bench_c.json 7 > gpurun_out/r02_convolve_bench_c.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_convolve_bench_c.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_c.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['best_known']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']); print(d['clocks']); print(d['device_api']['external']['worst_speedup'], d['device_api']['multiple']['worst_speedup'])"
du -sh gpurun_out
