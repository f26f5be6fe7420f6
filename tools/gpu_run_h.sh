#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/copylab 32 5 > gpurun_out/copylab_h.csv 2> gpurun_out/copylab_h.err; echo "rc=$?"
timeout 900 python tools/sweep_ctas.py gpurun_out/sweep_ctas_h.json 2>&1 | tail -20
