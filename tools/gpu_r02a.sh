#!/bin/bash
# Round 2, GPU call A: confirm the suite is green, then TIME everything round 1 built but never timed.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02a.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt 2>&1
nproc >> $O/r02a_smi.txt
timeout 60 python tools/quick_gpu_check.py --out $O/r02a_quick_check.json > $O/r02a_quick_check.txt 2>&1
echo "quick check exit $?" | tee -a $O/r02a_summary.txt
timeout 600 python -m pytest tests -m gpu -q -x > $O/r02a_pytest.log 2>&1
echo "gpu suite exit $?" | tee -a $O/r02a_summary.txt
tail -n 5 $O/r02a_pytest.log
# per-role timeline of the tensor-core kernels
timeout 200 python tools/tc_trace.py --json $O/r02a_tc_trace_S.json > $O/r02a_tc_trace_S.txt 2>&1
timeout 200 python tools/tc_trace.py --precision bf16x3 --json $O/r02a_tc_trace_S_bf16x3.json > $O/r02a_tc_trace_S_bf16x3.txt 2>&1
timeout 200 python tools/tc_trace.py --mask-sort --json $O/r02a_tc_trace_S_masksort.json > $O/r02a_tc_trace_S_masksort.txt 2>&1
head -n 40 $O/r02a_tc_trace_S.txt
# every prepared configuration per layer
timeout 500 python tools/autotune_conv.py --sweeps 1 --iters 20 --json $O/r02a_autotune.json > $O/r02a_autotune.txt 2>&1
tail -n 40 $O/r02a_autotune.txt
B="--steps 40 --warmup 10 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" > $O/r02a_bench_$name.json 2>$O/r02a_bench_$name.err; }
run L_S           A=1 python bench.py --workload L $B
run L_S_bf16x3    A=1 python bench.py --workload L $B --precision bf16x3
run L_S_masksort  MSMD_MASK_SORT=1 python bench.py --workload L $B
run L_S_bf16x3_masksort MSMD_MASK_SORT=1 python bench.py --workload L $B --precision bf16x3
run L_S_bf16x3_cps2 MSMD_TC_TUNE=cps=2 python bench.py --workload L $B --precision bf16x3
run L_S_bf16x3_cps2_masksort MSMD_TC_TUNE=cps=2 MSMD_MASK_SORT=1 python bench.py --workload L $B --precision bf16x3
run L_S_pdl       MSMD_LIB=msmdfusion_b200/_C/libmsmd_b200_pdl.so python bench.py --workload L $B
run L_L           A=1 python bench.py --workload L --profile L --steps 20 --warmup 5 --no-cpu-baseline
run L_L_bf16x3_masksort MSMD_MASK_SORT=1 python bench.py --workload L --profile L --steps 20 --warmup 5 --no-cpu-baseline --precision bf16x3
run LC_S          A=1 python bench.py --workload LC --steps 20 --warmup 5 --no-cpu-baseline --breakdown $O/r02a_breakdown_LC_S.json
run LC_S_bf16x3_masksort MSMD_MASK_SORT=1 python bench.py --workload LC --steps 20 --warmup 5 --no-cpu-baseline --precision bf16x3 --breakdown $O/r02a_breakdown_LC_S_bf16x3_masksort.json
run train         A=1 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --breakdown $O/r02a_breakdown_train.json
run train_wgradtc MSMD_WGRAD_TC=1 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --breakdown $O/r02a_breakdown_train_wgradtc.json
run train_bf16_wgradtc MSMD_WGRAD_TC=1 python bench.py --workload train --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline
for f in $O/r02a_bench_*.json; do
  echo "== $f"; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline') or {}
    print(round(d.get('value', 0), 2), 'scenes/s', round(d.get('ms_per_step', 0), 4), 'ms; e2e', round((d.get('e2e') or {}).get('value', 0), 2),
          '; conv ms', r.get('kernel_ms_per_step'), 'frac', r.get('frac'), 'launches', d.get('gpu_launches'))
except Exception as e:
    print('unparsed', e)
PY
done | tee -a $O/r02a_summary.txt
# launch list of one LC step (shares per kernel)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $O/r02a_launches_LC.csv \
  python bench.py --workload LC --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $O | tail -n 50
