#!/bin/bash
# which change moved the shape_from_shading LM results: the diagnostic matrix on three builds of the library
mkdir -p gpurun_out
{
  echo "== green build (db5ee85)"; OPTB_LIBOPT=build/variants/green/libOpt.so timeout 200 python tools/sfs_diag.py default generic no_iter_graph
  echo "== call-2 build (d999676)"; OPTB_LIBOPT=build/variants/call2/libOpt.so timeout 200 python tools/sfs_diag.py default generic no_resident per_step
  echo "== HEAD"; timeout 400 python tools/sfs_diag.py
} > gpurun_out/sfs_diag.txt 2> gpurun_out/sfs_diag.err
cut -c1-600 gpurun_out/sfs_diag.txt; tail -5 gpurun_out/sfs_diag.err
