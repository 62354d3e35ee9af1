#!/bin/bash
# Builds oracle/_ref/*.so from the reference's own sources WHERE THEY LIE (/root/reference), for pinning
# the oracle.  Outputs only under oracle/_ref/ (git-ignored, travels with gpurun).  No reference build
# system is run: each library is the shim in oracle/ plus the reference headers it includes.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference
[ -d "$REF" ] || { echo "ref_build: $REF not mounted, keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
CXX=/usr/bin/g++
FLAGS="-O2 -fopenmp -fPIC -shared -std=c++14 -w -I/usr/local/cuda/include"
if [ ! -f "$HERE/_ref/libref_iw.so" ] || [ "$HERE/ref_iw_shim.cpp" -nt "$HERE/_ref/libref_iw.so" ]; then
  $CXX $FLAGS -I"$REF/examples/image_warping/src" -o "$HERE/_ref/libref_iw.so" "$HERE/ref_iw_shim.cpp"
fi
if [ ! -f "$HERE/_ref/libref_arap.so" ] || [ "$HERE/ref_arap_shim.cpp" -nt "$HERE/_ref/libref_arap.so" ]; then
  $CXX $FLAGS -I"$REF/examples/arap_mesh_deformation/src" -o "$HERE/_ref/libref_arap.so" "$HERE/ref_arap_shim.cpp"
fi
if [ ! -f "$HERE/_ref/libref_sfs.so" ] || [ "$HERE/ref_sfs_shim.cpp" -nt "$HERE/_ref/libref_sfs.so" ]; then
  $CXX $FLAGS -I"$REF/examples/shape_from_shading/src" -o "$HERE/_ref/libref_sfs.so" "$HERE/ref_sfs_shim.cpp"
fi
if [ ! -f "$HERE/_ref/libref_poisson.so" ] || [ "$HERE/ref_poisson_shim.cpp" -nt "$HERE/_ref/libref_poisson.so" ]; then
  $CXX $FLAGS -I"$REF/examples/poisson_image_editing/src" -o "$HERE/_ref/libref_poisson.so" "$HERE/ref_poisson_shim.cpp"
fi
