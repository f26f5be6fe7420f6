#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest (new tests)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "32GiB or cli or in_place" 2>&1 | tee gpurun_out/pytest_y.log | tail -6
echo "=== sanitizer"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck_y.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_y.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck_y.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_y.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py > gpurun_out/sanitizer_synccheck_y.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/sanitizer_synccheck_y.log
