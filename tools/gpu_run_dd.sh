#!/bin/bash
mkdir -p gpurun_out
L=smfft_b200/lib
echo "=== pytest (real)"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "r2c or real or R2C or golden or c2r" 2>&1 | tail -3
echo "=== A/B scalar no-mirror (A) vs product (B)"; timeout 600 python tools/ab.py $L/libsmfft_a0.so $L/libsmfft.so gpurun_out/ab_mirror_c2r.json 2048,1024,4096
timeout 600 python tools/ab.py $L/libsmfft_a0.so $L/libsmfft.so gpurun_out/ab_mirror_c2r2.json 2048
