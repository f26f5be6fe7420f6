#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_m.log | tail -6
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_m.json 9 > gpurun_out/report_m.log 2>&1; echo "report rc=$?"; grep -E "^r2c" gpurun_out/report_m.log | cut -c1-420
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['clocks'])"
