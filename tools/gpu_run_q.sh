#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_aa.log | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_aa.json 2> gpurun_out/bench_aa.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_aa.err; python -c "
import json; d=json.load(open('gpurun_out/bench_aa.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['clocks'], d['cpu_baseline'])"
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_aa.json 9 > gpurun_out/report_aa.log 2>&1; echo "report rc=$?"
