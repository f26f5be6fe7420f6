#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2 1; do
  if [ $n -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_w_n$n.json 2> gpurun_out/bench_w_n$n.err
  else
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_w_n1.json 2> gpurun_out/bench_w_n1.err
  fi
  echo "N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_w_n$n.json')); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'])" 2>&1 | tail -1
done
