#!/bin/bash
# builds tools/tune_dual: the dual-lane shape sweep only (tools/tune.cu with -DTUNE_ONLY_DUAL); run as  tools/tune_dual 29 9 0 3
set -e
cd "$(dirname "$0")"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -I../include -I../smfft_b200/csrc -I. -DTUNE_ONLY_DUAL"
mkdir -p _build/dual
for f in tune tune_sizes_dual_a tune_sizes_dual_b tune_sizes_dual_c tune_sizes_dual_d tune_sizes_dual_e; do
  /usr/local/cuda/bin/nvcc $FLAGS -c $f.cu -o _build/dual/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ _build/dual/*.o -lcuda -o tune_dual
echo built tools/tune_dual
