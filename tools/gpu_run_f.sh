#!/bin/bash
mkdir -p gpurun_out
timeout 900 tools/tune 29 7 10 > gpurun_out/tune_f.csv 2> gpurun_out/tune_f.err; echo "rc=$?"; tail -2 gpurun_out/tune_f.err; wc -l gpurun_out/tune_f.csv
