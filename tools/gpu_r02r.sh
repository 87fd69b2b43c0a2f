#!/bin/bash
# Round 2, GPU call R: per-shape schedule dispatch + programmatic dependent launch of the persistent kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02r_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee $O/r02r_summary.txt
tail -n 4 $O/r02r_pytest_all.log | tee -a $O/r02r_summary.txt
cp $O/parity_abs_err.json $O/r02r_parity_abs_err.json 2>/dev/null
for pdl in 1 0; do
  MSMD_SB_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02r_bench_LC_S_pdl$pdl.json 2>$O/r02r_bench_LC_S_pdl$pdl.err
  MSMD_SB_PDL=$pdl timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02r_bench_L_S_pdl$pdl.json 2>$O/r02r_bench_L_S_pdl$pdl.err
done
MSMD_SB_VARIANT=1 timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02r_bench_L_S_tile.json 2>$O/r02r_bench_L_S_tile.err
timeout 600 python tools/lc_timeline.py --steps 1 > $O/r02r_lc_timeline.txt 2>&1
python - <<'PY' | tee -a gpurun_out/r02r_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02r_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'], 2), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 2),
              '; frac', r.get('frac'), '; kernel ms', r.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
grep "step 0" $O/r02r_lc_timeline.txt
