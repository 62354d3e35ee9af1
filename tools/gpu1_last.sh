#!/bin/bash
# last single-GPU call of the round (about four minutes of budget): the shape_from_shading fix first, then its tests, its config, a launch
# list and -- if the clock allows -- the headline bench line with the final library
mkdir -p gpurun_out
timeout 120 python tools/sfs_diag.py default no_head_in_loop > gpurun_out/sfs_diag_final.txt 2> gpurun_out/sfs_diag_final.err
timeout 400 python -m pytest tests/test_gpu_sfs_on_chip.py tests/test_gpu_engine_cross_check.py tests/test_gpu_generic_energies.py tests/test_golden_and_capi.py -q -m gpu --tb=short -rf 2>&1 | cut -c1-300 | tail -80 > gpurun_out/sfs_tests_final.txt
timeout 200 python -m pytest tests/test_gpu_fixtures_and_full_size.py tests/test_gpu_replan_params.py -q -m gpu -k "shape or replan" --tb=short -rf 2>&1 | cut -c1-300 | tail -30 >> gpurun_out/sfs_tests_final.txt
{
  echo "== default"; timeout 60 python tools/bench_configs.py sfs
  echo "== head not in the loop kernel"; OPTB_SFS_NO_HEAD_IN_LOOP=1 timeout 60 python tools/bench_configs.py sfs
} > gpurun_out/sfs_toggles_final.txt 2> gpurun_out/sfs_toggles_final.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r2_launches_sfs_tiles.csv python tools/bench_configs.py sfs > /dev/null 2>&1
timeout 150 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-500 gpurun_out/sfs_diag_final.txt; grep -E "passed|failed|^FAILED" gpurun_out/sfs_tests_final.txt; cut -c1-330 gpurun_out/sfs_toggles_final.txt; cut -c1-300 gpurun_out/bench_final.json
