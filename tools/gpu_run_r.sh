#!/bin/bash
mkdir -p gpurun_out
for v in "" _e5a _e5b ""; do
  SMFFT_LIB=$PWD/smfft_b200/lib/libsmfft$v.so timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('$v', round(d['value'],1), {k:v['ms'] for k,v in d['per_size'].items() if k in ('32r','32n','64r','4096r')}, d['clocks']['sm_mhz'])"
done
