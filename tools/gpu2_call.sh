#!/bin/bash
# One 2-GPU gpurun call (charged 2x): multi-rank parity tests, the collective loop kernel at the tile shape 8 GPUs will use for
# 2048^2 (1024^2 on 2 ranks -> 512x7 tiles, 148 CTAs), and bench lines with the parity_vs_n1 record.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu2_call.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_persistent.py -q -rxXs 2>&1 | tail -40 > gpurun_out/tests_2gpu.txt
{
  echo "== 1024^2 on 2 ranks, collective loop kernel (default when it fits)"
  timeout 200 $TR --master-port 29621 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
  echo "== 1024^2 on 2 ranks, per-step loop"
  OPTB_PERSISTENT=0 timeout 200 $TR --master-port 29622 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
  echo "== 2048^2 on 2 ranks (per-step loop: the slab does not fit on chip)"
  timeout 300 $TR --master-port 29623 bench.py --gpus 2 --steps 8 --warmup 3
  echo "== arap 500k on 1 and 2 ranks"
  timeout 200 python tools/bench_graph_multi.py 2>&1 | tail -3
  timeout 200 $TR --master-port 29624 tools/bench_graph_multi.py 2>&1 | tail -3
} > gpurun_out/bench_2gpu.txt 2>&1
tail -6 gpurun_out/tests_2gpu.txt; grep -E "^==|\"value\"" gpurun_out/bench_2gpu.txt | cut -c1-400
