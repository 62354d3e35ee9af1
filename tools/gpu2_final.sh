#!/bin/bash
# final 2-GPU pass: the multi-rank parity tests over NVLink and the headline at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_persistent.py -q -rxXs -k "2-" 2>&1 | tail -8 > gpurun_out/tests_2gpu_final.txt
{
  echo "== 2048^2 on 2 ranks"; timeout 240 $TR --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 4
  echo "== arap 500k on 2 ranks"; timeout 200 $TR --master-port 29742 tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
  echo "== rna 1M fp64 on 2 ranks"; timeout 240 $TR --master-port 29743 tools/bench_graph_multi.py --workload rna1m 2>&1 | tail -2
} > gpurun_out/bench_2gpu_final.txt 2>&1
tail -3 gpurun_out/tests_2gpu_final.txt; grep -E "^==|\"value\"|pcg_steps_per_s" gpurun_out/bench_2gpu_final.txt | cut -c1-200
