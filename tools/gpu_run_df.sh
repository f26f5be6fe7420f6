#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
echo "=== sustained bench: product vs 4096-natural-on-R16 variant, alternating"
for v in "" _r16 "" _r16; do
  SMFFT_LIB=$L/libsmfft$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-baselines --no-other-modes > gpurun_out/bench_sus2$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_sus2$v.json')); print('product$v', round(d['value'],1), round(d['ms_per_4GiB_batch'],4), {k:v['ms'] for k,v in d['per_size'].items() if k in ('32r','1024r','2048r','2048n','4096r','4096n')}, d['clocks']['sm_mhz'])"
done
