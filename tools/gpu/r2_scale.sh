#!/bin/bash
# multi-GPU pass: the bench as the driver launches it, plus the bare copy ceiling at the same rank count
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/r02_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac']); print(d['e2e']); print(d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/copy_peak.py gpurun_out/r02_copy_peak_n$N.json 2 2>&1 | tail -1
