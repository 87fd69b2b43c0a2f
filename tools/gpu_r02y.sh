#!/bin/bash
# Round 2, GPU call Y: final validation of HEAD -- smoke, the whole GPU suite, the driver's two bench commands.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02y_smoke.log 2>&1
echo "smoke exit $?" | tee $O/r02y_summary.txt
tail -n 1 $O/r02y_smoke.log | tee -a $O/r02y_summary.txt
timeout 900 python -m pytest tests -m gpu -q > $O/r02y_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee -a $O/r02y_summary.txt
tail -n 3 $O/r02y_pytest_all.log | tee -a $O/r02y_summary.txt
cp $O/parity_abs_err.json $O/r02y_parity_abs_err.json 2>/dev/null
timeout 900 python bench.py > $O/r02y_bench_default.json 2>$O/r02y_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02y_bench_reference.json 2>$O/r02y_bench_reference.err
python - <<'PY' | tee -a gpurun_out/r02y_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02y_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'], 3), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 3),
              '; frac', r.get('frac'), '; kernel ms', r.get('kernel_ms_per_step'), '; traffic', r.get('traffic'), '; cuda', (d.get('cuda_baseline') or {}).get('value'),
              '; cpu', (d.get('cpu_baseline') or {}).get('value'), '; launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
