#!/bin/bash
# Single-GPU measurement call of round 2: the whole -m gpu suite, the secondary configs, J^T J kernel-shape variants, the bench line, and
# the ncu evidence (launch list of the bench command + one full capture per dominant kernel) -> gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu1_profile_call.sh'
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -q -m gpu -rxXs 2>&1 | tail -120 > gpurun_out/tests.txt
{
  timeout 300 python tools/bench_configs.py iw512 sfs arap500k rna
  OPTB_SFS_NO_TILED=1 timeout 200 python tools/bench_configs.py sfs
} > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
timeout 300 python tools/bench_graph_multi.py --workload rna1m > gpurun_out/rna1m.json 2> gpurun_out/rna1m.err
{
  for v in th8s2 th8s3 th8s4 th16s3; do
    echo "== $v"; OPTB_LIBOPT=build/variants/$v/libOpt.so timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernels_ms'])"
  done
} > gpurun_out/variants.txt 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
# per-rank problem sizes of the 2/4/8-GPU runs on ONE GPU, no communication: what the per-step loop costs without the sync points
for sz in "2048 1024" "2048 512" "2048 256"; do set -- $sz; timeout 200 python tools/bench_slab_single.py $1 $2 >> gpurun_out/slab_single.jsonl 2>> gpurun_out/bench.err; done
OPTB_BENCH_SERIAL_E2E=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_serial_e2e.json 2>> gpurun_out/bench.err
# ncu: launch list of the bench command, then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/r2_launches_iw2048.csv python bench.py --steps 2 --warmup 3 --liters 40 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_iwApplyJTJTma -s 5 -c 1 -o gpurun_out/r2_jtj -f python bench.py --steps 1 --warmup 3 --liters 8 --no-cpu-baseline > gpurun_out/ncu_jtj.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcgStep2 -s 5 -c 1 -o gpurun_out/r2_step2 -f python bench.py --steps 1 --warmup 3 --liters 8 --no-cpu-baseline > gpurun_out/ncu_step2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sfsApplyJTJTiled -s 5 -c 1 -o gpurun_out/r2_sfs_jtj -f python tools/bench_configs.py sfs > gpurun_out/ncu_sfs.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_rnaApplyJTJ -s 5 -c 1 -o gpurun_out/r2_rna_jtj -f python tools/bench_graph_multi.py --workload rna1m > gpurun_out/ncu_rna.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_arapApplyJTJ -s 5 -c 1 -o gpurun_out/r2_arap_jtj -f python tools/bench_configs.py arap500k > gpurun_out/ncu_arap.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r2_launches_sfs.csv python tools/bench_configs.py sfs > /dev/null 2>&1
tail -4 gpurun_out/tests.txt; cat gpurun_out/variants.txt; cut -c1-200 gpurun_out/bench.json
