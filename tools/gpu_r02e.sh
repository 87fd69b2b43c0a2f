#!/bin/bash
# Round 2, GPU call E: bucketed exact FPS (first hardware run), LC timeline (host vs GPU), full suite.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fps" > $O/r02e_pytest_fps.log 2>&1
echo "fps tests exit $?" | tee $O/r02e_summary.txt
tail -n 6 $O/r02e_pytest_fps.log
timeout 200 python tools/fps_bench.py --json $O/r02e_fps_bench.json 2>&1 | tee $O/r02e_fps_bench.txt
timeout 200 python tools/lc_timeline.py --steps 2 --json $O/r02e_lc_timeline.json > $O/r02e_lc_timeline.txt 2>&1
cat $O/r02e_lc_timeline.txt | tail -n 45
MSMD_LC_OVERLAP=0 timeout 200 python tools/lc_timeline.py --steps 1 > $O/r02e_lc_timeline_nooverlap.txt 2>&1
B="--no-cpu-baseline --no-cuda-baseline"
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c --breakdown $O/r02e_breakdown_LC_S_sb.json > $O/r02e_bench_LC_S_sb.json 2>$O/r02e_bench_LC_S_sb.err
python - <<'PY' | tee -a gpurun_out/r02e_summary.txt
import json
d = json.loads(open('gpurun_out/r02e_bench_LC_S_sb.json').read().strip().splitlines()[-1])
print('LC', round(d['value'], 2), 'scenes/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 2))
PY
timeout 900 python -m pytest tests -m gpu -q > $O/r02e_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee -a $O/r02e_summary.txt
tail -n 6 $O/r02e_pytest_all.log
