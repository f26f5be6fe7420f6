#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== burst A/B product (A) vs register-direct MINB=8 default (B)"; timeout 300 python tools/ab.py $L/libsmfft.so $L/libsmfft_rd8.so gpurun_out/ab_rd8.json 1024 2>&1 | tail -1
echo "=== sustained bench: product vs register-direct (8 CTAs/SM) for 1024 natural, alternating"
for v in "" _rd8 "" _rd8; do
  SMFFT_LIB=$L/libsmfft$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-baselines --no-other-modes > gpurun_out/bench_sus5$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_sus5$v.json')); print('product$v', round(d['value'],1), round(d['ms_per_4GiB_batch'],4), {k:(v['ms'],v['ms_min']) for k,v in d['per_size'].items() if k in ('512r','1024r','1024n','2048r')}, d['clocks']['sm_mhz'])"
done
