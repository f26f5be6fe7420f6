#!/bin/bash
# builds tools/tune (shape sweep) for sm_100a; objects compile in parallel
set -e
cd "$(dirname "$0")"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -I../include -I../smfft_b200/csrc -I."
mkdir -p _build
for f in tune tune_sizes_a tune_sizes_b tune_sizes_c tune_sizes_d tune_sizes_real_a tune_sizes_real_b tune_sizes_real_c tune_sizes_reg_a tune_sizes_reg_b tune_sizes_reg_c tune_sizes_reg_d tune_sizes_reg_e; do
  /usr/local/cuda/bin/nvcc $FLAGS -c $f.cu -o _build/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ _build/*.o -o tune
echo built tools/tune
