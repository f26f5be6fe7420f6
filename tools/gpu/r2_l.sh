#!/bin/bash
# round 2: 8192 points, racecheck after the warp sync, bench with 8192 row
mkdir -p gpurun_out
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tee gpurun_out/r02_pytest_l.log | tail -5
echo "=== racecheck"; timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck_l.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_racecheck_l.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_l.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_l.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['traffic']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['clocks'], d['e2e']['value'], d['e2e']['frac']); print(d['other_modes']['c2c_8192']); print({k:v['ms'] for k,v in d['other_modes']['ct_multiple'].items()})"
