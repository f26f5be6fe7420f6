#!/bin/bash
mkdir -p gpurun_out
python tools/burst_timeline.py gpurun_out/r02_burst_timeline_product.json 80 2>&1 | tail -8
SMFFT_LIB=$PWD/smfft_b200/lib/libsmfft_ldg1.so python tools/burst_timeline.py gpurun_out/r02_burst_timeline_ldg1.json 80 2>&1 | grep reg_a
SMFFT_LIB=$PWD/smfft_b200/lib/libsmfft_ldst1.so python tools/burst_timeline.py gpurun_out/r02_burst_timeline_ldst1.json 80 2>&1 | grep reg_a
