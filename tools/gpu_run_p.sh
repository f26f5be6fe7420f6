#!/bin/bash
mkdir -p gpurun_out
timeout 1500 tools/tune 29 11 > gpurun_out/tune_v.csv 2> gpurun_out/tune_v.err; echo "rc=$?"; tail -2 gpurun_out/tune_v.err; wc -l gpurun_out/tune_v.csv
