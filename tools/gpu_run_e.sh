#!/bin/bash
mkdir -p gpurun_out
timeout 900 tools/copylab 32 5 > gpurun_out/copylab_g.csv 2> gpurun_out/copylab_g.err; echo "rc=$?"; tail -2 gpurun_out/copylab_g.err; wc -l gpurun_out/copylab_g.csv
