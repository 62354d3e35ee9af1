#!/bin/bash
# quick 2-GPU A/B of the collective loop kernel (1024^2 on 2 ranks = the 512x7 tile shape 8 GPUs use for 2048^2) + parity tests
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi_persistent.py tests/test_gpu_multi.py -q -rxXs -k "not arap and not rna" 2>&1 | tail -15 > gpurun_out/tests_${N}gpu_quick.txt
{
  echo "== 1024^2 on 2 ranks, collective loop kernel"
  timeout 200 $TR --master-port 29621 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
  echo "== 2048^2 on $N ranks"
  timeout 300 $TR --master-port 29623 bench.py --gpus $N --steps 8 --warmup 3
  echo "== arap 500k on $N ranks"
  timeout 200 $TR --master-port 29624 tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
  echo "== arap 500k on $N ranks, fused Step3"
  OPTB_ARAP_FUSION=1 timeout 200 $TR --master-port 29627 tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
} > gpurun_out/bench_${N}gpu_quick.txt 2>&1
tail -3 gpurun_out/tests_${N}gpu_quick.txt; grep -E "^==|\"value\"|pcg_steps_per_s|rror" gpurun_out/bench_${N}gpu_quick.txt | cut -c1-260
