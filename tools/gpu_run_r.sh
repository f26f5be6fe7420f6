#!/bin/bash
mkdir -p gpurun_out
echo "=== tune reg"; timeout 900 tools/tune 29 5 0 2 > gpurun_out/tune_reg.csv 2> gpurun_out/tune_reg.err; echo "rc=$?"; wc -l gpurun_out/tune_reg.csv; tail -3 gpurun_out/tune_reg.err
echo "=== ab real multiple: A = product, B = R32 real multiple"; timeout 600 python tools/ab.py smfft_b200/lib/libsmfft.so smfft_b200/lib/libsmfft_r32rm.so gpurun_out/ab_rmult.json 512,1024 2>&1 | tail -3
