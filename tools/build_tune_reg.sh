#!/bin/bash
# builds tools/tune_reg: the register-direct (IO_REG) shape sweep only; run as  TUNE_CARVEOUT=-1 tools/tune_reg 29 7 0 2
set -e
cd "$(dirname "$0")"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -I../include -I../smfft_b200/csrc -I. -DTUNE_ONLY_REG"
mkdir -p _build/reg
for f in tune tune_sizes_reg_a tune_sizes_reg_b tune_sizes_reg_c tune_sizes_reg_d tune_sizes_reg_e; do
  /usr/local/cuda/bin/nvcc $FLAGS -c $f.cu -o _build/reg/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ _build/reg/*.o -lcuda -o tune_reg
echo built tools/tune_reg
