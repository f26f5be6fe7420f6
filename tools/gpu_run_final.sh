#!/bin/bash
# last pass of the round: GPU tests + smoke + the default bench line + sanitizers on the final kernels
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_final.log | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; wc -l gpurun_out/bench_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac'], d['roofline']['traffic']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['clocks'], d['cpu_baseline']['value']); o=d['other_modes']; print({k:v['ms'] for k,v in o['r2c'].items()}); print({k:v['ms'] for k,v in o['c2r'].items()}); print({k:v['ms'] for k,v in o['ct_multiple'].items()}); print(d['baselines']['cufft_ms'])"
echo "=== sanitizer"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck_final.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_final.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck_final.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_final.log
