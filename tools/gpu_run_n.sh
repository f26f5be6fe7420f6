#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/real_variants.py gpurun_out/real_variants_n.json 2>&1 | tail -10
