#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 tools/tune 29 5 0 0 > gpurun_out/tune_small_$i.csv 2>/dev/null; echo "rc=$?"; done
