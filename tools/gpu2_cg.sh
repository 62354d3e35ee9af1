#!/bin/bash
# 2 GPUs over NVLink: the collective loop kernel at the 8-GPU tile shape (1024^2 on 2 ranks), two-reduction (default) vs opt-in single-reduction
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
{
  echo "== 1024^2 on 2 ranks, two-reduction kernel (default)"
  timeout 200 $TR --master-port 29721 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
  echo "== 1024^2 on 2 ranks, single-reduction kernel (OPTB_CG_LOOP=1)"
  OPTB_CG_LOOP=1 timeout 200 $TR --master-port 29722 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
} > gpurun_out/bench_2gpu_cg.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi_persistent.py -q -rxXs 2>&1 | tail -8 > gpurun_out/tests_2gpu_cg.txt
tail -3 gpurun_out/tests_2gpu_cg.txt; grep -E "^==|\"value\"" gpurun_out/bench_2gpu_cg.txt | cut -c1-200
