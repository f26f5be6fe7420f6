#!/bin/bash
# two-pass transforms: parity tests, sanitizers on the small target
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_pass or beyond" 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_target.py > gpurun_out/r02_sanitizer_${tool}_twopass.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/r02_sanitizer_${tool}_twopass.log; done
