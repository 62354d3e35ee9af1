#!/bin/bash
# Builds alternative libOpt.so files with different compile-time shapes of the TMA-tiled J^T J kernel (energy_image_warping.cu:
# IW_TH rows per tile, IW_STAGES pipeline depth, IW_STRIP rows per thread; 256 threads = 64 columns x IW_TH/IW_STRIP strips) into
# build/variants/<name>/libOpt.so for A/B timing on the GPU box:  OPTB_LIBOPT=build/variants/th16s3/libOpt.so python bench.py ...
# Measured (round 2, fused kernel in the bench loop): th16s2 (2 CTAs/SM) 59.9 us, th8s2 61.1, th8s3 59.3, th8s4 59.5, th16s3 (1 CTA/SM) 55.4.
set -e
cd "$(dirname "$0")/../opt_b200/csrc"
build() {  # name flags
  local out=../../build/variants/$1
  mkdir -p $out/obj
  make -s -j8 OUT=$out OBJ=$out/obj EXTRA_NVFLAGS="$2" $out/libOpt.so
  echo "built $out/libOpt.so ($2)"
}
for v in "$@"; do
  case $v in
    th8s2) build th8s2 "-DIW_TH=8 -DIW_STAGES=2 -DIW_STRIP=2";;
    th8s3) build th8s3 "-DIW_TH=8 -DIW_STAGES=3 -DIW_STRIP=2";;
    th8s4) build th8s4 "-DIW_TH=8 -DIW_STAGES=4 -DIW_STRIP=2";;
    th16s3) build th16s3 "-DIW_TH=16 -DIW_STAGES=3 -DIW_STRIP=4";;
    th16s2) build th16s2 "-DIW_TH=16 -DIW_STAGES=2 -DIW_STRIP=4";;
    th16s4) build th16s4 "-DIW_TH=16 -DIW_STAGES=4 -DIW_STRIP=4";;
    th32s2) build th32s2 "-DIW_TH=32 -DIW_STAGES=2 -DIW_STRIP=8";;
    th8s6) build th8s6 "-DIW_TH=8 -DIW_STAGES=6 -DIW_STRIP=2";;
    *) echo "unknown variant $v"; exit 1;;
  esac
done
