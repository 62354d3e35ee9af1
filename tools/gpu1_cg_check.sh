#!/bin/bash
# single-GPU validation of the single-reduction loop kernel: its parity tests, the multi-rank tests with ranks sharing the GPU, and timing
# on one GPU (512^2: on-chip standard loop vs single-reduction loop)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_image_warping.py tests/test_gpu_multi.py tests/test_gpu_multi_persistent.py tests/test_gpu_resident_tile.py tests/test_gpu_replan_params.py tests/test_gpu_example_steps.py -q -rxXs 2>&1 | tail -40 > gpurun_out/tests_cg.txt
{
  echo "== 512^2 default (row-slab on-chip loop)"; timeout 120 python tools/bench_configs.py iw512
  echo "== 512^2 single-reduction loop"; OPTB_CG_LOOP=1 timeout 120 python tools/bench_configs.py iw512
  echo "== 512^2 two-reduction tile loop 64x32"; OPTB_RESIDENT_TILE=64x32 OPTB_PERSISTENT=1 timeout 120 python tools/bench_configs.py iw512
} > gpurun_out/cg_single.txt 2>&1
tail -6 gpurun_out/tests_cg.txt; grep -E "^==|pcg_loop_us" gpurun_out/cg_single.txt | cut -c1-330
