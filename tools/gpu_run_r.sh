#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
echo "=== ab multiple: A = product, B = R32 multiple"; timeout 600 python tools/ab.py smfft_b200/lib/libsmfft.so smfft_b200/lib/libsmfft_r32m.so gpurun_out/ab_mult.json 512,1024,4096 2>&1 | tail -4
echo "=== tune real e12"; timeout 900 tools/tune 29 5 12 1 > gpurun_out/tune_real_e12.csv 2> gpurun_out/tune_real_e12.err; echo "rc=$?"; cut -d, -f1-10,20,22 gpurun_out/tune_real_e12.csv | sort -t, -k12 -n | head -30
