#!/bin/bash
# 16384-point C2C, R = 32 plan [32,32,16]: both stagings, with and without the spare half-tile buffer
mkdir -p gpurun_out
for v in "" _t14nosplit; do
  lib=smfft_b200/lib/libsmfft$v.so
  echo "=== $lib"
  SMFFT_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "beyond or 16384" 2>&1 | tail -2
  SMFFT_LIB=$PWD/$lib timeout 600 python tools/ab_16384.py | tee gpurun_out/r02_ab_16384$v.json | cut -c1-400
done
