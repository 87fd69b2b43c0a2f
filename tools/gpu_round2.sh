#!/bin/bash
# First GPU call of round 2 (everything written at the end of round 1 without GPU time), in priority order.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh'
# Outputs land in gpurun_out/r02a_*; each step has its own timeout so that one hang cannot eat the call.
set -u
mkdir -p gpurun_out
O=gpurun_out
# 0. torch-free smoke of the late-round-1 kernels (2 s): if this fails, read it before anything else
timeout 60 python tools/quick_gpu_check.py --out $O/r02a_quick_check.json > $O/r02a_quick_check.txt 2>&1
echo "quick check exit $?" | tee -a $O/r02a_summary.txt
# 1. the forward-path suite first (must stay green), then the new backward / mask-sort tests on their own
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_train_gpu.py > $O/r02a_pytest_forward.log 2>&1
echo "forward suite exit $?" | tee -a $O/r02a_summary.txt
timeout 600 python -m pytest tests/test_zz_train_gpu.py -q > $O/r02a_pytest_train_masksort.log 2>&1
echo "train + mask-sort suite exit $?" | tee -a $O/r02a_summary.txt
tail -n 30 $O/r02a_pytest_train_masksort.log
# 1b. where does the per-chunk time go?  per-role timeline of the tensor-core kernels (csrc/tc_trace.cuh)
timeout 300 python tools/tc_trace.py --json $O/r02a_tc_trace_S.json > $O/r02a_tc_trace_S.txt 2>&1
timeout 300 python tools/tc_trace.py --precision bf16x3 --json $O/r02a_tc_trace_S_bf16x3.json > $O/r02a_tc_trace_S_bf16x3.txt 2>&1
timeout 300 python tools/tc_trace.py --mask-sort --json $O/r02a_tc_trace_S_masksort.json > $O/r02a_tc_trace_S_masksort.txt 2>&1
timeout 300 python tools/tc_trace.py --sweeps 10 --json $O/r02a_tc_trace_L.json > $O/r02a_tc_trace_L.txt 2>&1
head -n 30 $O/r02a_tc_trace_S.txt
# 1c. every prepared configuration of the conv kernel, per layer of both profiles (the table that decides defaults)
timeout 600 python tools/autotune_conv.py --json $O/r02a_autotune.json > $O/r02a_autotune.txt 2>&1
tail -n 45 $O/r02a_autotune.txt
# 2. bench lines: default, mask-sorted, 16-bit operand modes, LC, LC mask-sorted, train step (fp32 / bf16 / tc wgrad)
timeout 300 python bench.py --steps 100 --warmup 30 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_S_bf16x3.json 2>$O/r02a_bench_S_bf16x3.err
MSMD_MASK_SORT=1 timeout 300 python bench.py --steps 100 --warmup 30 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_S_bf16x3_masksort.json 2>&1
MSMD_TC16_VARIANT=3 timeout 300 python bench.py --steps 100 --warmup 30 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_S_bf16x3_v3.json 2>&1   # only meaningful if the quick check's variant-3 lines say ok
MSMD_TC_TUNE=cps=2 timeout 300 python bench.py --steps 100 --warmup 30 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_S_bf16x3_cps2.json 2>&1
MSMD_TC_TUNE=cps=2 MSMD_MASK_SORT=1 timeout 300 python bench.py --steps 100 --warmup 30 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_S_bf16x3_cps2_masksort.json 2>&1
timeout 300 python bench.py --profile L --steps 50 --warmup 10 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_L_bf16x3.json 2>&1
timeout 300 python bench.py --workload LC --steps 30 --warmup 10 --precision bf16x3 --no-cpu-baseline > $O/r02a_bench_LC_bf16x3.json 2>&1
# programmatic dependent launch of the conv chain (debug build, same ABI): A/B against the default line
python -m msmdfusion_b200.build --pdl > $O/r02a_build_pdl.log 2>&1
MSMD_LIB=msmdfusion_b200/_C/libmsmd_b200_pdl.so timeout 300 python bench.py --steps 100 --warmup 30 --no-cpu-baseline > $O/r02a_bench_S_pdl.json 2>&1
for b in 2 4; do timeout 300 python bench.py --steps 60 --warmup 20 --scenes-per-step $b --no-cpu-baseline > $O/r02a_bench_S_batch$b.json 2>&1; done
MSMD_WGRAD_TC=1 timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline --breakdown $O/r02a_breakdown_train_wgradtc.json > $O/r02a_bench_train_wgradtc.json 2>&1
MSMD_WGRAD_TC=1 timeout 300 python bench.py --workload train --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02a_bench_train_bf16_wgradtc.json 2>&1
timeout 300 python bench.py --steps 100 --warmup 30 > $O/r02a_bench_S.json 2>$O/r02a_bench_S.err
MSMD_MASK_SORT=1 timeout 300 python bench.py --steps 100 --warmup 30 --no-cpu-baseline > $O/r02a_bench_S_masksort.json 2>$O/r02a_bench_S_masksort.err
MSMD_MASK_SORT=1 timeout 300 python bench.py --profile L --steps 50 --warmup 10 --no-cpu-baseline > $O/r02a_bench_L_masksort.json 2>&1
timeout 300 python bench.py --profile L --steps 50 --warmup 10 --no-cpu-baseline > $O/r02a_bench_L.json 2>&1
timeout 300 python bench.py --workload LC --steps 30 --warmup 10 --no-cpu-baseline > $O/r02a_bench_LC.json 2>&1
MSMD_MASK_SORT=1 timeout 300 python bench.py --workload LC --steps 30 --warmup 10 --no-cpu-baseline > $O/r02a_bench_LC_masksort.json 2>&1
timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline --breakdown $O/r02a_breakdown_train.json > $O/r02a_bench_train.json 2>$O/r02a_bench_train.err
for f in S S_masksort S_bf16x3 S_bf16x3_masksort S_bf16x3_cps2 S_bf16x3_cps2_masksort S_bf16x3_v3 S_pdl S_batch2 S_batch4 L L_masksort L_bf16x3 LC LC_masksort LC_bf16x3 train train_wgradtc train_bf16_wgradtc; do
  echo "== $f"; tail -c 600 $O/r02a_bench_$f.json | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
except Exception as e: print('unparsed', e)"
done | tee -a $O/r02a_summary.txt
# 3. ncu: launch list of the train step, full capture of the wgrad kernel
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 --csv --log-file $O/r02a_launches_train.csv \
  python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spconv_wgrad -s 16 -c 16 -o $O/r02a_prof_wgrad \
  python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $O | tail -n 20
