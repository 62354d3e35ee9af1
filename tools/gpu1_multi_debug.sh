#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port"
{
  echo "== 2 ranks 128x96 GN, single-reduction kernel"; OPTB_SHARE_GPU=1 timeout 200 $TR 29711 tests/multi_gpu_check.py 128 96 gaussNewtonGPU 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | head -60
  echo "== 2 ranks 128x64 GN, single-reduction kernel"; OPTB_SHARE_GPU=1 timeout 200 $TR 29712 tests/multi_gpu_check.py 128 64 gaussNewtonGPU 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | head -40
} > gpurun_out/multi_debug.txt 2>&1
cat gpurun_out/multi_debug.txt | cut -c1-250 | head -70
