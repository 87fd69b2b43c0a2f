#!/bin/bash
# Round 2, GPU call D: elect.sync MMA issue + overlapped image-side schedule; full suite with the absolute bound.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02d_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee $O/r02d_summary.txt
tail -n 8 $O/r02d_pytest_all.log
timeout 300 python tools/tc_trace.py --precision bf16x3c --json $O/r02d_tc_trace_S_sb.json > $O/r02d_tc_trace_S_sb.txt 2>&1
cat $O/r02d_tc_trace_S_sb.txt | tail -n 15
B="--no-cpu-baseline --no-cuda-baseline"
timeout 300 python bench.py --workload L --steps 40 --warmup 10 $B --precision bf16x3c > $O/r02d_bench_L_S_sb.json 2>$O/r02d_bench_L_S_sb.err
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c --breakdown $O/r02d_breakdown_LC_S_sb.json > $O/r02d_bench_LC_S_sb.json 2>$O/r02d_bench_LC_S_sb.err
MSMD_LC_OVERLAP=0 timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c > $O/r02d_bench_LC_S_sb_nooverlap.json 2>$O/r02d_bench_LC_S_sb_nooverlap.err
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B > $O/r02d_bench_LC_S_tf32.json 2>$O/r02d_bench_LC_S_tf32.err
timeout 300 python bench.py --workload LC --profile L --steps 10 --warmup 3 $B --precision bf16x3c > $O/r02d_bench_LC_L_sb.json 2>$O/r02d_bench_LC_L_sb.err
for f in $O/r02d_bench_*.json; do
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
done | tee -a $O/r02d_summary.txt
tail -3 $O/r02d_bench_LC_S_sb.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02d_launches_L_sb.csv \
  python bench.py --workload L --steps 2 --warmup 3 $B --precision bf16x3c > $O/r02d_launches_L_sb.log 2>&1
tail -2 $O/r02d_launches_L_sb.log
