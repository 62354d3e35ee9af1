#!/bin/bash
# 8-GPU call (charged 8x: keep it short): the headline scaling point with the parity_vs_n1 record, the per-step loop for comparison,
# and BASELINE config 5 (robust_nonrigid_alignment fp64, 8 ranks, domain-partitioned).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
{
  echo "== 2048^2 on $N ranks (default: collective loop kernel when every slab fits)"
  timeout 240 $TR --master-port 29631 bench.py --gpus $N --steps 20 --warmup 5
  echo "== 2048^2 on $N ranks, per-step loop forced"
  OPTB_PERSISTENT=0 timeout 240 $TR --master-port 29632 bench.py --gpus $N --steps 20 --warmup 5
  echo "== rna 1M fp64 LM on $N ranks"
  timeout 240 $TR --master-port 29633 tools/bench_graph_multi.py --workload rna1m 2>&1 | tail -2
  echo "== arap 500k on $N ranks"
  timeout 200 $TR --master-port 29634 tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
} > gpurun_out/bench_${N}gpu.txt 2>&1
grep -E "^==|\"value\"|pcg_steps_per_s|rror" gpurun_out/bench_${N}gpu.txt | cut -c1-260
