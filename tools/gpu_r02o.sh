#!/bin/bash
# Round 2, GPU call O: persistent kernel v2 (row-contiguous epilogue, tile masks / weighted shares, mask-sorted option)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02o_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee $O/r02o_summary.txt
tail -n 4 $O/r02o_pytest_all.log | tee -a $O/r02o_summary.txt
timeout 600 python tools/sb_bench.py --lc --json $O/r02o_sb_bench_LC.json > $O/r02o_sb_bench_LC.txt 2>&1
tail -n 3 $O/r02o_sb_bench_LC.txt
timeout 300 python tools/tc_trace.py --precision bf16x3c --sb-variant 0 --dump-cta --only "128->128 k27,32->32 k27,128->128 k3,64->64 k27" --json $O/r02o_tc_trace_S_sbp.json > $O/r02o_tc_trace_S_sbp.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02o_bench_LC_S.json 2>$O/r02o_bench_LC_S.err
MSMD_MASK_SORT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02o_bench_LC_S_masksort.json 2>$O/r02o_bench_LC_S_masksort.err
timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02o_bench_L_S.json 2>$O/r02o_bench_L_S.err
MSMD_MASK_SORT=1 timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02o_bench_L_S_masksort.json 2>$O/r02o_bench_L_S_masksort.err
python - <<'PY' | tee -a gpurun_out/r02o_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02o_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'], 2), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 2),
              '; frac', r.get('frac'), '; kernel ms', r.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
tail -3 $O/r02o_bench_LC_S.err
