#!/bin/bash
# ncu --set full of the device-API kernels (final engines) next to the reference's, N = 1024 both orders + N = 32
mkdir -p gpurun_out /tmp/ncu
for cfg in "compat external 1024 0 1" "compat external 1024 0 0" "reference external 1024 0 1" "reference external 1024 0 0" "compat multiple 1024 0 1" "compat multiple 1024 0 0" "reference multiple 1024 0 1" "reference multiple 1024 0 0" "compat multiple 32 0 1" "reference multiple 32 0 1"; do set -- $cfg
timeout 600 ncu --set full --clock-control none --import-source on -k regex:SMFFT_DIT -s 2 -c 1 -f -o /tmp/ncu/r02f_$1_$2_n$3_i$4_r$5 python tools/compat_target.py $1 $2 $3 $4 $5 > /dev/null 2>&1; echo "ncu $cfg rc=$?"
done
python tools/ncu_summarize.py gpurun_out/r02_ncu_summary_compat_final.md /tmp/ncu/r02f_*.ncu-rep > /dev/null 2>&1; echo "summarize rc=$?"
