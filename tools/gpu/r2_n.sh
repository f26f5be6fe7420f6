#!/bin/bash
mkdir -p gpurun_out
L=$PWD/smfft_b200/lib
echo "=== 8192-point shapes: product [R32, 2 stages, 1 CTA] | 3 stages | R16 (4 passes, 512 threads)"
timeout 300 python tools/ab_libs.py gpurun_out/r02_ab_8192_shapes.json 8192 $L/libsmfft.so $L/libsmfft_t13s3.so $L/libsmfft_t13b4.so 2>&1 | tail -3
timeout 300 python tools/ab_libs.py gpurun_out/r02_ab_8192_shapes_burst.json 8192 --burst 20 $L/libsmfft.so $L/libsmfft_t13s3.so $L/libsmfft_t13b4.so 2>&1 | tail -3
