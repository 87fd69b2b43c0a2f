#!/bin/bash
# Round 2, GPU call B: where does the conv kernel's per-chunk time go (ncu source-level + per-role trace),
# first run of the reference CUDA baseline.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/tc_trace.py --json $O/r02b_tc_trace_S.json > $O/r02b_tc_trace_S.txt 2>&1
timeout 300 python tools/tc_trace.py --precision bf16x3 --json $O/r02b_tc_trace_S_bf16x3.json > $O/r02b_tc_trace_S_bf16x3.txt 2>&1
cat $O/r02b_tc_trace_S.txt | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_fwd -s 21 -c 21 -f -o $O/r02b_prof_conv_S \
  python tools/prof_conv.py > $O/r02b_prof_conv_S.log 2>&1
tail -3 $O/r02b_prof_conv_S.log
timeout 300 python bench.py --workload L --steps 20 --warmup 5 --no-cpu-baseline > $O/r02b_bench_L_S.json 2>$O/r02b_bench_L_S.err
timeout 600 python bench.py --workload LC --steps 20 --warmup 5 > $O/r02b_bench_LC_S.json 2>$O/r02b_bench_LC_S.err
tail -c 1500 $O/r02b_bench_LC_S.json; tail -5 $O/r02b_bench_LC_S.err
timeout 300 python bench.py --impl reference --workload LC --steps 2 --warmup 0 > $O/r02b_bench_LC_S_reference.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02b_launches_LC.csv \
  python bench.py --workload LC --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/r02b_launches_LC.log 2>&1
ls -la $O | grep r02b
