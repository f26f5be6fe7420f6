#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== staging option probe (real transforms)"; timeout 300 python tools/io_probe.py 2>&1 | tail -4
echo "=== sustained bench: scalar build vs packed-everywhere build (C2C step), alternating"
for v in a0 a2 a0 a2; do
  SMFFT_LIB=$L/libsmfft_$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-baselines --no-other-modes > gpurun_out/bench_sus_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_sus_$v.json')); print('$v', round(d['value'],1), round(d['ms_per_4GiB_batch'],4), {k:v['ms'] for k,v in d['per_size'].items() if k in ('32r','1024r','2048r','2048n','4096r','4096n')}, d['clocks']['sm_mhz'])"
done
