#!/bin/bash
# Compiles the REFERENCE's own test programs (tests/minimal, tests/minimal_graph_only,
# tests/create_delete_cycle main.cpp) UNCHANGED, from where they lie under /root/reference, against this
# repository's include/Opt.h and opt_b200/lib/libOpt.so -- the API-conformance check of SURVEY.md 2 #24.
# Outputs only under tests/_refbin/ (git-ignored; travels to the GPU box with gpurun).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
REF=/root/reference/tests
[ -d "$REF" ] || { echo "build_reference_tests: $REF not mounted, keeping prebuilt tests/_refbin"; exit 0; }
mkdir -p "$HERE/_refbin"
for t in minimal minimal_graph_only create_delete_cycle; do
  out="$HERE/_refbin/$t"
  if [ ! -f "$out" ] || [ "$ROOT/opt_b200/lib/libOpt.so" -nt "$out" ]; then
    /usr/bin/g++ -O1 -w -std=c++14 -I"$ROOT/include" -I/usr/local/cuda/include "$REF/$t/main.cpp" -o "$out" \
      -L"$ROOT/opt_b200/lib" -Wl,-rpath,'$ORIGIN/../../opt_b200/lib' -lOpt -L/usr/local/cuda/lib64 -lcudart
  fi
done
