#!/bin/bash
# Round 2, GPU call I: GMA stage geometry on the geometry stream (rulebooks ahead of the convolutions); two-strand schedule.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02i_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee $O/r02i_summary.txt
tail -n 6 $O/r02i_pytest_all.log
timeout 200 python tools/lc_timeline.py --steps 2 --json $O/r02i_lc_timeline.json > $O/r02i_lc_timeline.txt 2>&1
timeout 200 python tools/lc_hostprofile.py --steps 20 --top 30 > $O/r02i_lc_hostprofile.txt 2>&1
B="--no-cpu-baseline --no-cuda-baseline"
timeout 300 python bench.py --workload L --steps 40 --warmup 10 $B --precision bf16x3c > $O/r02i_bench_L_S_sb.json 2>$O/r02i_bench_L_S_sb.err
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c --breakdown $O/r02i_breakdown_LC_S_sb.json > $O/r02i_bench_LC_S_sb.json 2>$O/r02i_bench_LC_S_sb.err
MSMD_GMA_NATIVE=0 timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c > $O/r02i_bench_LC_S_sb_nonative.json 2>$O/r02i_bench_LC_S_sb_nonative.err
for f in $O/r02i_bench_*.json; do
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
done | tee -a $O/r02i_summary.txt
tail -3 $O/r02i_bench_LC_S_sb.err
