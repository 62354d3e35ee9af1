#!/bin/bash
# third targeted single-GPU check of the shape_from_shading tile kernels (head inside the loop kernel, tagged tile-edge words, one-pass
# precompute): all their tests with complete output, the config under the new toggles, the launch list of a run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sfs_on_chip.py tests/test_gpu_engine_cross_check.py -q -m gpu --tb=short -rf 2>&1 | cut -c1-400 > gpurun_out/sfs_tests3.txt
timeout 600 python -m pytest tests/test_gpu_generic_energies.py tests/test_gpu_fixtures_and_full_size.py tests/test_golden_and_capi.py tests/test_gpu_replan_params.py -q -m gpu -k "shape or sfs or replan" --tb=short -rf 2>&1 | cut -c1-400 | tail -40 >> gpurun_out/sfs_tests3.txt
{
  echo "== default"; timeout 120 python tools/bench_configs.py sfs
  echo "== head not in the loop kernel"; OPTB_SFS_NO_HEAD_IN_LOOP=1 timeout 120 python tools/bench_configs.py sfs
  echo "== two-launch precompute"; OPTB_SFS_NO_FUSED_PRECOMPUTE=1 timeout 120 python tools/bench_configs.py sfs
  echo "== 1024 threads"; OPTB_SFS_LOOP_NT=1024 timeout 120 python tools/bench_configs.py sfs
  echo "== 832 threads"; OPTB_SFS_LOOP_NT=832 timeout 120 python tools/bench_configs.py sfs
} > gpurun_out/sfs_toggles3.txt 2> gpurun_out/sfs_toggles3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r2_launches_sfs_tiles.csv python tools/bench_configs.py sfs > /dev/null 2>&1
grep -c . gpurun_out/sfs_tests3.txt; grep -E "passed|failed" gpurun_out/sfs_tests3.txt; grep -E "^FAILED|^E  " gpurun_out/sfs_tests3.txt | head -40; cut -c1-420 gpurun_out/sfs_toggles3.txt; tail -5 gpurun_out/sfs_toggles3.err
