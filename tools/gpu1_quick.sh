#!/bin/bash
# quick single-GPU check after kernel changes: full -m gpu suite, SFS / RNA configs, J^T J kernel-shape variants
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -q -m gpu -rxXs 2>&1 | tail -60 > gpurun_out/tests.txt
{
  timeout 300 python tools/bench_configs.py sfs rna
} > gpurun_out/configs_quick.jsonl 2> gpurun_out/configs.err
timeout 300 python tools/bench_graph_multi.py --workload rna1m > gpurun_out/rna1m.json 2> gpurun_out/rna1m.err
{
  for v in th16s3 th16s4 th32s2 th8s6; do
    echo "== $v"; OPTB_LIBOPT=build/variants/$v/libOpt.so timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernels_ms'])"
  done
} > gpurun_out/variants.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r2_launches_sfs.csv python tools/bench_configs.py sfs > /dev/null 2>&1
tail -4 gpurun_out/tests.txt; cat gpurun_out/configs_quick.jsonl | cut -c1-330; cut -c1-300 gpurun_out/rna1m.json; cat gpurun_out/variants.txt
