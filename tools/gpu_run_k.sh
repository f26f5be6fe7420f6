#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_k.log | tail -6
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']); print(d['clocks'])"
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_k.json 7 > gpurun_out/report_k.log 2>&1; echo "report rc=$?"
