#!/bin/bash
# Round 2, GPU call K: bf16x3c as the library default; smoke(); FPS CTA-width A/B; default bench line (what the driver runs).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02k_smoke.log 2>&1
echo "smoke exit $?" | tee $O/r02k_summary.txt
tail -n 2 $O/r02k_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02k_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee -a $O/r02k_summary.txt
tail -n 5 $O/r02k_pytest_all.log
for w in 256 512 1024; do
  MSMD_FPS_THREADS=$w timeout 200 python tools/lc_timeline.py --steps 1 > $O/r02k_lc_timeline_fps$w.txt 2>&1
  echo "fps width $w:" | tee -a $O/r02k_summary.txt
  grep "  fps" $O/r02k_lc_timeline_fps$w.txt | tee -a $O/r02k_summary.txt
  grep "step 0" $O/r02k_lc_timeline_fps$w.txt | tee -a $O/r02k_summary.txt
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02k_bench_default.json 2>$O/r02k_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02k_bench_reference.json 2>$O/r02k_bench_reference.err
python - <<'PY' | tee -a gpurun_out/r02k_summary.txt
import json
for f in ('r02k_bench_default', 'r02k_bench_reference'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f, round(d['value'], 3), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 3), '; frac', r.get('frac'),
              '; cuda_baseline', (d.get('cuda_baseline') or {}).get('value'), '; cpu', (d.get('cpu_baseline') or {}).get('value'))
        print('   config', d['config']['workload'][:80])
    except Exception as e:
        print(f, 'unparsed', e)
PY
tail -3 $O/r02k_bench_default.err
