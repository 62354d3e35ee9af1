#!/bin/bash
# One single-GPU gpurun call: the whole -m gpu suite (with xfail/xpass report), two ranks time-slicing ONE GPU through the
# multi-rank paths (per-step loop and the collective one-kernel loop), the secondary configs, and the bench line.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -q -m gpu -rxXs 2>&1 | tail -150 > gpurun_out/tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port"
{
  echo "== shared GPU, per-step loop"; OPTB_SHARE_GPU=1 OPTB_PERSISTENT=0 timeout 300 $TR 29611 tests/multi_gpu_check.py 128 96 gaussNewtonGPU 2>&1 | tail -8
  echo "== shared GPU, collective loop kernel"; OPTB_SHARE_GPU=1 OPTB_TEST_MAX_LAUNCHES=12 timeout 300 $TR 29612 tests/multi_gpu_check.py 128 96 gaussNewtonGPU 2>&1 | tail -8
  echo "== shared GPU, arap"; OPTB_SHARE_GPU=1 timeout 300 $TR 29614 tests/multi_gpu_check.py arap 48 40 gaussNewtonGPU 2>&1 | tail -6
  echo "== shared GPU, rna fp64 LM"; OPTB_SHARE_GPU=1 timeout 300 $TR 29615 tests/multi_gpu_check.py rna 40 48 LMGPU f64 2>&1 | tail -6
  echo "== shared GPU, collective loop kernel, 64x8 tiles"; OPTB_SHARE_GPU=1 OPTB_TEST_MAX_LAUNCHES=12 OPTB_RESIDENT_TILE=64x8 timeout 300 $TR 29613 tests/multi_gpu_check.py 256 130 gaussNewtonGPU 2>&1 | tail -8
} > gpurun_out/shared_gpu.txt 2>&1
{
  timeout 300 python tools/bench_configs.py iw512 sfs arap500k
  OPTB_GENERIC=1 timeout 200 python tools/bench_configs.py sfs
  timeout 300 python tools/bench_configs.py rna
  OPTB_GENERIC=1 timeout 300 python tools/bench_configs.py rna
} > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/cross_n_divergence.py > gpurun_out/divergence.json 2> gpurun_out/divergence.err
tail -5 gpurun_out/tests.txt; cat gpurun_out/shared_gpu.txt | grep -E "==|multi_gpu_check|launches|Error|error" | head -20; cut -c1-300 gpurun_out/bench.json
