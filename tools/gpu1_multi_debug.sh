#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port"
{
  echo "== 2 ranks 128x96 GN, single-reduction kernel"; OPTB_SHARE_GPU=1 timeout 200 $TR 29711 tests/multi_gpu_check.py 128 96 gaussNewtonGPU 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -8
  echo "== 2 ranks 128x96 GN, two-reduction kernel"; OPTB_CG_LOOP=0 OPTB_SHARE_GPU=1 timeout 200 $TR 29712 tests/multi_gpu_check.py 128 96 gaussNewtonGPU 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -8
  echo "== 2 ranks 256x130 64x8 tiles, single-reduction kernel"; OPTB_RESIDENT_TILE=64x8 OPTB_SHARE_GPU=1 timeout 200 $TR 29713 tests/multi_gpu_check.py 256 130 gaussNewtonGPU 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -8
  echo "== 4 ranks 192x64 GN, single-reduction kernel"; OPTB_SHARE_GPU=1 timeout 300 $TR4 29714 tests/multi_gpu_check.py 192 64 gaussNewtonGPU 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -8
  echo "== 2 ranks 256x130 LM"; OPTB_SHARE_GPU=1 timeout 200 $TR 29715 tests/multi_gpu_check.py 256 130 LMGPU 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -8
  echo "== 2 ranks arap"; OPTB_SHARE_GPU=1 timeout 200 $TR 29716 tests/multi_gpu_check.py arap 48 40 gaussNewtonGPU 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -6
} > gpurun_out/multi_debug.txt 2>&1
cat gpurun_out/multi_debug.txt | cut -c1-250
