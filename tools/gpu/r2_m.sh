#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== 8192-point shapes: product [R32, 2 stages, 1 CTA] | 3 stages | 1 stage x 2 CTAs | R16 (4 passes)"
python tools/ab_libs.py gpurun_out/r02_ab_8192_shapes.json 8192 $L/libsmfft.so $L/libsmfft_t13s3.so $L/libsmfft_t13c2.so $L/libsmfft_t13b4.so 2>&1 | tail -3
python tools/ab_libs.py gpurun_out/r02_ab_8192_shapes_burst.json 8192 --burst 20 $L/libsmfft.so $L/libsmfft_t13s3.so $L/libsmfft_t13c2.so $L/libsmfft_t13b4.so 2>&1 | tail -3
echo "=== pytest io 4/5"; python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2c_vs_oracle or in_place_every" 2>&1 | tail -2
echo "=== select probe (one-warp register-direct B shapes at 512 / 1024)"
python tools/select_probe.py gpurun_out/r02_select_probe_m.json 15 2>&1 | grep -v "^\"\|^{\|^}" | grep -v "0/0:0.0000"
