#!/bin/bash
# 4-GPU call: the 4-rank parity tests over NVLink, the headline at N=4 with the parity_vs_n1 record, graph configs at N=4 (BASELINE config 4: arap 1->4)
N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_persistent.py -q -rxXs -k "4-" 2>&1 | tail -12 > gpurun_out/tests_4gpu.txt
{
  echo "== 2048^2 on 4 ranks"
  timeout 240 $TR --master-port 29641 bench.py --gpus 4 --steps 12 --warmup 4
  echo "== arap 500k on 4 ranks"
  timeout 200 $TR --master-port 29642 tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
  echo "== rna 1M fp64 LM on 4 ranks"
  timeout 240 $TR --master-port 29643 tools/bench_graph_multi.py --workload rna1m 2>&1 | tail -2
} > gpurun_out/bench_4gpu.txt 2>&1
tail -3 gpurun_out/tests_4gpu.txt; grep -E "^==|\"value\"|pcg_steps_per_s|rror" gpurun_out/bench_4gpu.txt | cut -c1-260
