#!/bin/bash
# Round 2, GPU call S: A/B of the geometry-stream priority and of the helper thread that issues the compression block
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overlapped or voxel_space" > $O/r02s_pytest_lc.log 2>&1
echo "lc tests exit $?" | tee $O/r02s_summary.txt
tail -n 3 $O/r02s_pytest_lc.log | tee -a $O/r02s_summary.txt
for prio in 1 0; do for thr in 1 0; do
  MSMD_GEOM_PRIORITY=$prio MSMD_LC_HOST_THREAD=$thr timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02s_bench_LC_S_prio${prio}_thr${thr}.json 2>$O/r02s_bench_LC_S_prio${prio}_thr${thr}.err
done; done
for prio in 1 0; do
  MSMD_GEOM_PRIORITY=$prio timeout 600 python bench.py --workload L --steps 30 --warmup 5 --no-cuda-baseline --no-cpu-baseline > $O/r02s_bench_L_S_prio${prio}.json 2>$O/r02s_bench_L_S_prio${prio}.err
done
timeout 600 python tools/lc_timeline.py --steps 1 > $O/r02s_lc_timeline.txt 2>&1
python - <<'PY' | tee -a gpurun_out/r02s_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02s_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'], 2), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 2),
              '; frac', r.get('frac'), '; kernel ms', r.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
grep "step 0" $O/r02s_lc_timeline.txt
