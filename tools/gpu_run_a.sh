#!/bin/bash
# first on-box pass: sanity, shape sweep, parity tests, golden fixtures from the reference kernels, bench
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
echo "=== tune"; timeout 900 tools/tune 29 5 > gpurun_out/tune_a.csv 2> gpurun_out/tune_a.err; echo "tune rc=$?"; tail -3 gpurun_out/tune_a.err; wc -l gpurun_out/tune_a.csv
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee gpurun_out/pytest_a.log | tail -25
echo "=== golden"; timeout 300 python tests/golden/make_golden.py gpurun_out/golden 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_a.err; cut -c1-1500 gpurun_out/bench_a.json
