#!/bin/bash
# dual-lane shape sweep (tools/tune_dual: product shapes next to the dual-lane candidates, same box)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 tools/tune_dual 29 9 0 3 > gpurun_out/tune_dual_a.csv 2> gpurun_out/tune_dual_a.err; echo "rc=$?"; tail -3 gpurun_out/tune_dual_a.err; wc -l gpurun_out/tune_dual_a.csv
