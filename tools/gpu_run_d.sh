#!/bin/bash
mkdir -p gpurun_out
echo "=== copylab"; timeout 1200 tools/copylab 32 5 > gpurun_out/copylab_d.csv 2> gpurun_out/copylab_d.err; echo "rc=$?"; tail -2 gpurun_out/copylab_d.err; wc -l gpurun_out/copylab_d.csv
echo "=== ncu cufft"
for n in 256 1024; do
timeout 600 ncu --set full --clock-control none --import-source on -s 1 -c 1 -f -o gpurun_out/prof_cufft_n$n python tools/cufft_target.py $n > gpurun_out/ncu_cufft_$n.log 2>&1; echo "ncu cufft $n rc=$?"
done
