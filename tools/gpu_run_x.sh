#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_x.log | tail -4
echo "=== ncu full at the 4 GiB batch (traffic for the roofline)"
SMFFT_NCU_LOG2_POINTS=29 timeout 900 ncu --set full --clock-control none --import-source on -k regex:smfft_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_x_c2c_n1024_4GiB python tools/ncu_target.py c2c 1024 1 > gpurun_out/ncu_full_x.log 2>&1; echo "rc=$?"
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print(d['value'], d['ms_per_4GiB_batch'], d['roofline']['frac']); print({k:v['ms'] for k,v in d['per_size'].items()}); print(d['e2e']['value'], d['clocks'])"
