#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "=== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_o_n2.json 2> gpurun_out/bench_o_n2.err; echo "rc=$?"; tail -3 gpurun_out/bench_o_n2.err; cut -c1-300 gpurun_out/bench_o_n2.json
echo "=== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_o_n1.json 2> gpurun_out/bench_o_n1.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_o_n1.json
echo "=== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
