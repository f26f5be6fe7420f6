#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== pytest (real + golden)"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "r2c or real or R2C or golden or c2r" 2>&1 | tail -2
echo "=== A/B scalar baseline (A) vs product (B)"; timeout 600 python tools/ab.py $L/libsmfft_a0.so $L/libsmfft.so gpurun_out/ab_dh1.json 512,1024,2048
echo "=== A/B product (A) vs tail-barrier experiment (B)"; timeout 600 python tools/ab.py $L/libsmfft.so $L/libsmfft_tb.so gpurun_out/ab_dh2.json 256,4096
