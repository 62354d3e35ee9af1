#!/bin/bash
# second targeted single-GPU check of the shape_from_shading tile kernels: all their tests, thread-count / fold toggles, one ncu --set full
# capture of the whole-loop kernel and the launch list of a run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sfs_on_chip.py tests/test_gpu_engine_cross_check.py -q -m gpu -rxXs 2>&1 | tail -60 > gpurun_out/sfs_tests2.txt
timeout 600 python -m pytest tests/test_gpu_generic_energies.py tests/test_gpu_fixtures_and_full_size.py tests/test_golden_and_capi.py tests/test_gpu_replan_params.py -q -m gpu -k "shape or sfs or replan" -rxXs 2>&1 | tail -25 >> gpurun_out/sfs_tests2.txt
{
  echo "== default"; timeout 120 python tools/bench_configs.py sfs
  echo "== 832 threads"; OPTB_SFS_LOOP_NT=832 timeout 120 python tools/bench_configs.py sfs
  echo "== 768 threads"; OPTB_SFS_LOOP_NT=768 timeout 120 python tools/bench_configs.py sfs
  echo "== update not folded"; OPTB_SFS_NO_FUSED_UPDATE=1 timeout 120 python tools/bench_configs.py sfs
  echo "== model cost and update not folded"; OPTB_SFS_NO_FUSED_UPDATE=1 OPTB_SFS_NO_FUSED_MODEL=1 timeout 120 python tools/bench_configs.py sfs
} > gpurun_out/sfs_toggles2.txt 2> gpurun_out/sfs_toggles2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2_launches_sfs_tiles.csv python tools/bench_configs.py sfs > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sfsLoopResident -s 3 -c 1 -o gpurun_out/r2_sfs_loop python tools/bench_configs.py sfs > gpurun_out/ncu_sfs_loop.log 2>&1
tail -30 gpurun_out/sfs_tests2.txt; cut -c1-420 gpurun_out/sfs_toggles2.txt; tail -5 gpurun_out/sfs_toggles2.err; ls -la gpurun_out/r2_sfs_loop.ncu-rep; tail -3 gpurun_out/ncu_sfs_loop.log
