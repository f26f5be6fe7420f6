#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== pytest repeated path"; python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "repeated or multiple" 2>&1 | tail -2
echo "=== A/B FFT_multiple: A = through the tile every repetition, B = product (natural order chained in registers)"
timeout 900 python tools/ab_multiple.py $L/libsmfft_mulsmem.so $L/libsmfft.so gpurun_out/r02_ab_multiple_in_registers.json 32,64,128,256,512,1024,2048,4096 2>&1 | tail -17
