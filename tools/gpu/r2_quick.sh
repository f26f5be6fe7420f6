#!/bin/bash
# quick regression pass: GPU tests + smoke + the default bench line
mkdir -p gpurun_out
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tee gpurun_out/r02_pytest_quick.log | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_quick.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['e2e']['frac'], d['clocks']); print(d['device_api']['external']['worst_speedup'], d['device_api']['multiple']['worst_speedup']); print({k:v.get('ours_vs_cufft_steady') for k,v in d['baselines']['steady_state_per_kernel'].items() if isinstance(v,dict)})"
