#!/bin/bash
mkdir -p gpurun_out
echo "=== e2e chunk probe"; timeout 600 python tools/e2e_probe.py 2>&1 | tail -8
