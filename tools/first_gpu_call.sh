#!/bin/bash
# One gpurun call that answers everything ROUND2_CHECKLIST.md asks of a single B200:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# Results land in gpurun_out/ (merged back): tests.txt (incl. which xfail-marked tests XPASSed), configs.jsonl, bench.json.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rxX 2>&1 | tail -60 > gpurun_out/tests.txt
{
  timeout 300 python tools/bench_configs.py iw512 sfs arap500k
  timeout 200 python tools/bench_configs.py sfs_hand
  timeout 300 python tools/bench_configs.py rna rna_hand
  OPTB_PERSISTENT=0 timeout 120 python tools/bench_configs.py iw512
  OPTB_PERSISTENT=1 OPTB_RESIDENT_TILE=64x32 timeout 120 python tools/bench_configs.py iw512
} > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/tests.txt; grep -c config gpurun_out/configs.jsonl; cut -c1-200 gpurun_out/bench.json
