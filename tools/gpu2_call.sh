#!/bin/bash
# One N-GPU gpurun call (charged N x): multi-rank parity tests over real NVLink, the collective loop kernel at the tile shape 8 GPUs use
# for 2048^2 (1024^2 on 2 ranks -> 512x7 tiles, 148 CTAs), bench lines with the parity_vs_n1 record, and the graph-domain configs.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu2_call.sh 2'
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_persistent.py -q -rxXs 2>&1 | tail -40 > gpurun_out/tests_${N}gpu.txt
{
  if [ "$N" = "2" ]; then
    echo "== 1024^2 on 2 ranks, collective loop kernel (default when it fits)"
    timeout 200 $TR --master-port 29621 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
    echo "== 1024^2 on 2 ranks, per-step loop"
    OPTB_PERSISTENT=0 timeout 200 $TR --master-port 29622 bench.py --gpus 2 --size 1024 --steps 6 --warmup 3
  fi
  echo "== 2048^2 on $N ranks"
  timeout 300 $TR --master-port 29623 bench.py --gpus $N --steps 8 --warmup 3
  echo "== 2048^2 on $N ranks, per-step loop forced"
  OPTB_PERSISTENT=0 timeout 300 $TR --master-port 29625 bench.py --gpus $N --steps 8 --warmup 3
  echo "== arap 500k on 1 and $N ranks"
  timeout 200 python tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
  timeout 200 $TR --master-port 29624 tools/bench_graph_multi.py --workload arap500k 2>&1 | tail -2
  echo "== rna 1M fp64 LM on 1 and $N ranks"
  timeout 300 python tools/bench_graph_multi.py --workload rna1m 2>&1 | tail -2
  timeout 300 $TR --master-port 29626 tools/bench_graph_multi.py --workload rna1m 2>&1 | tail -2
} > gpurun_out/bench_${N}gpu.txt 2>&1
tail -6 gpurun_out/tests_${N}gpu.txt; grep -E "^==|\"value\"|pcg_steps_per_s|Error|error" gpurun_out/bench_${N}gpu.txt | cut -c1-300
