#!/bin/bash
# A/B of the arithmetic flavours (scalar / product selection / packed add-sub everywhere) + GPU tests on the new product
mkdir -p gpurun_out /tmp/ncu
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
L=smfft_b200/lib
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tee gpurun_out/pytest_db.log | tail -4
echo "=== A/B scalar (A) vs product (B)"; timeout 600 python tools/ab.py $L/libsmfft_a0.so $L/libsmfft.so gpurun_out/ab_scalar_vs_product.json 32,64,128,256,512,1024,2048,4096
echo "=== A/B scalar (A) vs packed everywhere (B)"; timeout 600 python tools/ab.py $L/libsmfft_a0.so $L/libsmfft_a2.so gpurun_out/ab_scalar_vs_packed_all.json 32,64,128,256,512,1024,2048,4096
echo "=== ncu: dual-lane R2C 4096 reals and C2C 1024 (tune_dual, only_e, first matching launch)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:BlockCfg<11, 4, 2, 0, 1, 0, .*1>, 1, 0, 2, 1, 3, 1>' -s 3 -c 1 -f -o /tmp/ncu/prof_dual_r2c_e11 tools/tune_dual 27 1 11 3 > gpurun_out/ncu_dual_r2c.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:BlockCfg<10, 4, 4, 0, 1, 0, .*1>, 0, 0, 2, 1, 3, 1>' -s 3 -c 1 -f -o /tmp/ncu/prof_dual_c2c_e10 tools/tune_dual 27 1 10 3 > gpurun_out/ncu_dual_c2c.log 2>&1; echo "rc=$?"
python tools/ncu_summarize.py gpurun_out/ncu_summary_dual.md /tmp/ncu/prof_dual_*.ncu-rep > gpurun_out/ncu_summarize_dual.log 2>&1; echo "summarize rc=$?"
for f in /tmp/ncu/prof_dual_r2c_e11.ncu-rep; do [ -f $f ] && cp $f gpurun_out/; done
du -sh gpurun_out
