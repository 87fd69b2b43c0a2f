#!/bin/bash
# Round 2, GPU call M: the persistent work-balanced schedule of the split-operand kernel -- parity suite, per-layer A/B
# against the tile-per-CTA schedule, LC / L bench lines.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02m_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee $O/r02m_summary.txt
tail -n 4 $O/r02m_pytest_all.log | tee -a $O/r02m_summary.txt
cp $O/parity_abs_err.json $O/r02m_parity_abs_err.json 2>/dev/null
timeout 600 python tools/sb_bench.py --lc --json $O/r02m_sb_bench_LC.json > $O/r02m_sb_bench_LC.txt 2>&1
tail -n 40 $O/r02m_sb_bench_LC.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02m_bench_LC_S.json 2>$O/r02m_bench_LC_S.err
MSMD_SB_VARIANT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02m_bench_LC_S_tile.json 2>$O/r02m_bench_LC_S_tile.err
timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02m_bench_L_S.json 2>$O/r02m_bench_L_S.err
MSMD_SB_VARIANT=1 timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02m_bench_L_S_tile.json 2>$O/r02m_bench_L_S_tile.err
python - <<'PY' | tee -a gpurun_out/r02m_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02m_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'], 2), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 2),
              '; frac', r.get('frac'), '; kernel ms', r.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
tail -3 $O/r02m_bench_LC_S.err
