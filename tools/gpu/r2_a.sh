#!/bin/bash
# round 2, first GPU pass: baseline state of the tests, the compat-vs-reference table, same-protocol baselines, copy ceiling
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tee gpurun_out/r02_pytest_a.log | tail -4
echo "=== compat bench"; timeout 900 python tools/compat_bench.py gpurun_out/r02_compat_bench_a.json 7 > gpurun_out/r02_compat_bench_a.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_compat_bench_a.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_a.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['clocks']); print('cufft', {k:v['ms'] for k,v in d['baselines']['cufft_ms'].items() if isinstance(v,dict)}); print('ref', {k:v['ms'] for k,v in d['baselines']['reference_sm100a_ms'].items() if isinstance(v,dict)})"
echo "=== copy peak"; timeout 600 python tools/copy_peak.py gpurun_out/r02_copy_peak_n1.json 2 2>&1 | tail -2
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -25 > gpurun_out/r02_lscpu.txt; numactl -H >> gpurun_out/r02_lscpu.txt 2>&1
du -sh gpurun_out
