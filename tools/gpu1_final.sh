#!/bin/bash
# final single-GPU pass of the round: whole -m gpu suite, secondary configs, the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rxXs 2>&1 | tail -60 > gpurun_out/tests.txt
timeout 300 python tools/bench_configs.py iw512 sfs arap500k rna > gpurun_out/configs_final.jsonl 2> gpurun_out/configs.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
tail -4 gpurun_out/tests.txt; cut -c1-330 gpurun_out/configs_final.jsonl; cut -c1-400 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_reference.json
