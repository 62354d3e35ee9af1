#!/bin/bash
# Builds alternative libOpt.so files with different compile-time shapes of the TMA-tiled J^T J kernel (energy_image_warping.cu:
# IW_TH rows per tile, IW_STAGES pipeline depth, IW_STRIP rows per thread; 256 threads = 64 columns x IW_TH/IW_STRIP strips) into
# build/variants/<name>/libOpt.so for A/B timing on the GPU box:  OPTB_LIBOPT=build/variants/th8s2/libOpt.so python bench.py ...
set -e
cd "$(dirname "$0")/../opt_b200/csrc"
build() {  # name flags
  local out=../../build/variants/$1
  mkdir -p $out/obj
  make -s -j8 OUT=$out OBJ=$out/obj EXTRA_NVFLAGS="$2" $out/libOpt.so
  echo "built $out/libOpt.so ($2)"
}
build th8s2 "-DIW_TH=8 -DIW_STAGES=2 -DIW_STRIP=2"
build th8s3 "-DIW_TH=8 -DIW_STAGES=3 -DIW_STRIP=2"
build th8s4 "-DIW_TH=8 -DIW_STAGES=4 -DIW_STRIP=2"
build th16s3 "-DIW_TH=16 -DIW_STAGES=3 -DIW_STRIP=4"
