#!/bin/bash
# round 2, fourth GPU pass: alternates + first-use selection
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "=== pytest (selection, io 4/5)"; timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "selection or c2c_vs_oracle or in_place" 2>&1 | tee gpurun_out/r02_pytest_d.log | tail -6
echo "=== select probe"; timeout 900 python tools/select_probe.py gpurun_out/r02_select_probe_d.json 10 > gpurun_out/r02_select_probe_d.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r02_select_probe_d.log
du -sh gpurun_out
