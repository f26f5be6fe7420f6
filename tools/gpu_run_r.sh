#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest (real transforms)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "r2c or c2r or real or golden or host" 2>&1 | tail -3
echo "=== ab new vs old R2C tail (A = new, B = old tail + R16)"; timeout 600 python tools/ab.py smfft_b200/lib/libsmfft.so smfft_b200/lib/libsmfft_nor32.so gpurun_out/ab_r2c.json 32,64,128,256,512,1024,2048,4096 2>&1 | tail -9
echo "=== tune real"; timeout 900 tools/tune 29 5 0 1 > gpurun_out/tune_real.csv 2> gpurun_out/tune_real.err; echo "rc=$?"; wc -l gpurun_out/tune_real.csv
