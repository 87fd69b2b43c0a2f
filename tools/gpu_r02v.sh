#!/bin/bash
# Round 2, GPU call V: the shape-static half of the compression block as a CUDA graph -- parity tests of the LC path, A/B
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02v_pytest_lc.log 2>&1
echo "gpu suite exit $?" | tee $O/r02v_summary.txt
tail -n 3 $O/r02v_pytest_lc.log | tee -a $O/r02v_summary.txt
B="--no-cpu-baseline --no-cuda-baseline"
for g in 1 0; do for thr in 0 1; do
  if [ $g = 0 ] && [ $thr = 0 ]; then continue; fi
  MSMD_COMPRESS_GRAPH=$g MSMD_LC_HOST_THREAD=$thr timeout 300 python bench.py --steps 20 --warmup 5 $B > $O/r02v_bench_LC_S_graph${g}_thr${thr}.json 2>$O/r02v_bench_LC_S_graph${g}_thr${thr}.err
done; done
MSMD_LC_HOST_THREAD=0 timeout 300 python tools/lc_timeline.py --steps 1 > $O/r02v_lc_timeline.txt 2>&1
python - <<'PY' | tee -a gpurun_out/r02v_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02v_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['value'], 2), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 2))
    except Exception as e:
        print(f, 'unparsed', e)
PY
grep "step 0" $O/r02v_lc_timeline.txt
tail -5 $O/r02v_bench_LC_S_graph1_thr0.err
