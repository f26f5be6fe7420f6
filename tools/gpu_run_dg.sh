#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== pytest (real + golden)"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "r2c or real or R2C or golden or c2r" 2>&1 | tail -2
echo "=== sustained bench: product vs 512/1024-natural-on-R16+packed variant, alternating"
for v in "" _x16 "" _x16; do
  SMFFT_LIB=$L/libsmfft$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-baselines --no-other-modes > gpurun_out/bench_sus3$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_sus3$v.json')); print('product$v', round(d['value'],1), round(d['ms_per_4GiB_batch'],4), {k:v['ms'] for k,v in d['per_size'].items() if k in ('256r','512r','512n','1024r','1024n','2048r','4096r','4096n')}, d['clocks']['sm_mhz'])"
done
echo "=== A/B scalar no-mirror (A) vs product (B)"; timeout 600 python tools/ab.py $L/libsmfft_a0.so $L/libsmfft.so gpurun_out/ab_final.json 32,64,128,256,512,1024,2048,4096
