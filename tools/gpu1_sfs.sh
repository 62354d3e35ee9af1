#!/bin/bash
# targeted single-GPU check of the shape_from_shading tile kernels (fused LM head, whole-loop kernel, cost / model cost folding):
# their tests, the config under every toggle, the launch list of one run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sfs_on_chip.py tests/test_gpu_engine_cross_check.py tests/test_gpu_replan_params.py -q -m gpu -x -rxXs 2>&1 | tail -40 > gpurun_out/sfs_tests.txt
timeout 600 python -m pytest tests/test_gpu_generic_energies.py tests/test_gpu_fixtures_and_full_size.py tests/test_golden_and_capi.py -q -m gpu -k "shape or sfs" -rxXs 2>&1 | tail -25 >> gpurun_out/sfs_tests.txt
{
  echo "== default"; timeout 120 python tools/bench_configs.py sfs
  echo "== no model-cost folding"; OPTB_SFS_NO_FUSED_MODEL=1 timeout 120 python tools/bench_configs.py sfs
  echo "== no cost folding"; OPTB_SFS_NO_FUSED_COST=1 timeout 120 python tools/bench_configs.py sfs
  echo "== no fused head"; OPTB_SFS_NO_FUSED_INIT=1 timeout 120 python tools/bench_configs.py sfs
  echo "== no on-chip loop"; OPTB_SFS_NO_RESIDENT=1 timeout 120 python tools/bench_configs.py sfs
  echo "== none of them"; OPTB_SFS_NO_RESIDENT=1 OPTB_SFS_NO_FUSED_INIT=1 OPTB_SFS_NO_FUSED_COST=1 timeout 120 python tools/bench_configs.py sfs
  echo "== no iteration graph"; OPTB_NO_ITER_GRAPH=1 timeout 120 python tools/bench_configs.py sfs
} > gpurun_out/sfs_toggles.txt 2> gpurun_out/sfs_toggles.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_launches_sfs_tiles.csv python tools/bench_configs.py sfs > /dev/null 2>&1
cat gpurun_out/sfs_tests.txt | tail -30; cut -c1-420 gpurun_out/sfs_toggles.txt; tail -5 gpurun_out/sfs_toggles.err
