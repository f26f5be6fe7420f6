#!/bin/bash
mkdir -p gpurun_out
echo "=== sustained (default build)"; timeout 900 python tools/sustained.py gpurun_out/sustained_l_default.json 150 2>&1 | tail -50
echo "=== sustained (packed f32x2 build)"; SMFFT_LIB=$PWD/smfft_b200/lib/libsmfft_packed.so timeout 900 python tools/sustained.py gpurun_out/sustained_l_packed.json 150 2>&1 | tail -50
