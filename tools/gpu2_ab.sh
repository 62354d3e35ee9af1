#!/bin/bash
# 2 GPUs: per-step loop at 2048^2 with the 3-stage (default) and 2-stage J^T J kernel shapes, with and without the iteration graph
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
{
  echo "== 3 stages (default)"; timeout 200 $TR --master-port 29731 bench.py --gpus 2 --steps 8 --warmup 3
  echo "== 3 stages, no iteration graph"; OPTB_NO_ITER_GRAPH=1 timeout 200 $TR --master-port 29732 bench.py --gpus 2 --steps 8 --warmup 3
  echo "== 2 stages"; OPTB_LIBOPT=$PWD/build/variants/th16s2/libOpt.so timeout 200 $TR --master-port 29733 bench.py --gpus 2 --steps 8 --warmup 3
} > gpurun_out/bench_2gpu_ab.txt 2>&1
grep -E "^==|\"value\"" gpurun_out/bench_2gpu_ab.txt | cut -c1-160
