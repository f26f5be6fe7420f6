#!/bin/bash
# ncu --set full: cuFFT's 512-point kernel vs our register-direct A / B and the TMA instance (where do the cycles go?)
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vector_fft -s 1 -c 1 -f -o /tmp/ncu/r02_o_cufft_512 python tools/cufft_target.py 512 > gpurun_out/r02_o_cufft.log 2>&1; echo "cufft rc=$?"
for io in 0 4 5; do
SMFFT_IO=$io timeout 600 ncu --set full --clock-control none --import-source on -k regex:smfft_tile_kernel -s 2 -c 1 -f -o /tmp/ncu/r02_o_ours_512_io$io python tools/ncu_target.py c2c 512 1 > gpurun_out/r02_o_ours_$io.log 2>&1; echo "ours io=$io rc=$?"
done
python tools/ncu_summarize.py gpurun_out/r02_ncu_summary_512_o.md /tmp/ncu/r02_o_*.ncu-rep > gpurun_out/r02_o_summarize.log 2>&1; echo "summarize rc=$?"
for f in /tmp/ncu/r02_o_*.ncu-rep; do ncu -i $f --page details --csv 2>/dev/null | grep -E "Warp Cycles Per Issued|Eligible Warps|Issued Warp|No Eligible|Theoretical Occupancy|Achieved Occupancy|Registers Per|L1/TEX Hit|Mem Busy|Max Bandwidth|Duration|Shared Memory Configuration|Driver Shared|Dynamic Shared|Block Limit" | cut -d, -f5,12-15 > gpurun_out/$(basename $f .ncu-rep)_details.txt; done
ls -la gpurun_out | tail -8
