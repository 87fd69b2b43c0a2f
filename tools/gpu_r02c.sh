#!/bin/bash
# Round 2, GPU call C: first hardware run of the split-operand conv kernel (csrc/spconv_sb.cu).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_zz_train_gpu.py -q -x -k "split_operand" > $O/r02c_pytest_sb.log 2>&1
echo "sb tests exit $?" | tee $O/r02c_summary.txt
tail -n 15 $O/r02c_pytest_sb.log
timeout 400 python tools/sb_bench.py --json $O/r02c_sb_bench_S.json > $O/r02c_sb_bench_S.txt 2>&1
cat $O/r02c_sb_bench_S.txt | tail -n 25
timeout 400 python tools/sb_bench.py --lc --json $O/r02c_sb_bench_LC.json > $O/r02c_sb_bench_LC.txt 2>&1
tail -n 30 $O/r02c_sb_bench_LC.txt
timeout 300 python tools/tc_trace.py --precision bf16x3c --json $O/r02c_tc_trace_S_sb.json > $O/r02c_tc_trace_S_sb.txt 2>&1
cat $O/r02c_tc_trace_S_sb.txt | tail -n 16
B="--no-cpu-baseline --no-cuda-baseline"
timeout 300 python bench.py --workload L --steps 40 --warmup 10 $B --precision bf16x3c > $O/r02c_bench_L_S_sb.json 2>$O/r02c_bench_L_S_sb.err
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c --breakdown $O/r02c_breakdown_LC_S_sb.json > $O/r02c_bench_LC_S_sb.json 2>$O/r02c_bench_LC_S_sb.err
timeout 300 python bench.py --workload L --profile L --steps 20 --warmup 5 $B --precision bf16x3c > $O/r02c_bench_L_L_sb.json 2>$O/r02c_bench_L_L_sb.err
for f in $O/r02c_bench_*.json; do
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
done | tee -a $O/r02c_summary.txt
tail -3 $O/r02c_bench_L_S_sb.err
timeout 900 python -m pytest tests -m gpu -q > $O/r02c_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee -a $O/r02c_summary.txt
tail -n 12 $O/r02c_pytest_all.log
