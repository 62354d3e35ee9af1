#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1"
{ echo "== 2048^2 on 4 ranks"; timeout 240 $TR --master-port 29751 bench.py --gpus 4 --steps 12 --warmup 4; } > gpurun_out/bench_4gpu_final2.txt 2>&1
grep -E "^==|\"value\"" gpurun_out/bench_4gpu_final2.txt | cut -c1-200
