#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv
timeout 1500 tools/tune 29 11 > gpurun_out/tune_i.csv 2> gpurun_out/tune_i.err; echo "rc=$?"; tail -2 gpurun_out/tune_i.err; wc -l gpurun_out/tune_i.csv
