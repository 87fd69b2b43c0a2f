#!/bin/bash
# Round 2, GPU call J: hash-path modality split, compression on the side stream; parity of the whole suite with
# bf16x3c as the process default; ncu --set full of the LC conv launches (traffic / tensor-pipe evidence).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02j_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee $O/r02j_summary.txt
tail -n 5 $O/r02j_pytest_all.log
cp $O/parity_abs_err.json $O/r02j_parity_abs_err_tf32x3.json
MSMD_CONV_PRECISION=bf16x3c timeout 900 python -m pytest tests -m gpu -q > $O/r02j_pytest_all_bf16x3c.log 2>&1
echo "gpu suite (bf16x3c default) exit $?" | tee -a $O/r02j_summary.txt
tail -n 12 $O/r02j_pytest_all_bf16x3c.log
cp $O/parity_abs_err.json $O/r02j_parity_abs_err_bf16x3c.json
timeout 200 python tools/lc_timeline.py --steps 1 --json $O/r02j_lc_timeline.json > $O/r02j_lc_timeline.txt 2>&1
B="--no-cpu-baseline --no-cuda-baseline"
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --precision bf16x3c --breakdown $O/r02j_breakdown_LC_S_sb.json > $O/r02j_bench_LC_S_sb.json 2>$O/r02j_bench_LC_S_sb.err
timeout 300 python bench.py --workload LC --profile L --steps 10 --warmup 3 $B --precision bf16x3c > $O/r02j_bench_LC_L_sb.json 2>$O/r02j_bench_LC_L_sb.err
timeout 300 python bench.py --workload L --steps 40 --warmup 10 $B --precision bf16x3c > $O/r02j_bench_L_S_sb.json 2>$O/r02j_bench_L_S_sb.err
for f in $O/r02j_bench_*.json; do
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
done | tee -a $O/r02j_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_fwd_sb -s 37 -c 37 -f -o $O/r02j_prof_conv_LC_S_sb \
  python tools/prof_conv.py --lc --precision bf16x3c > $O/r02j_prof_conv_LC.log 2>&1
tail -2 $O/r02j_prof_conv_LC.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fps_cluster -c 1 -f -o $O/r02j_prof_fps \
  python tools/prof_conv.py --lc --precision bf16x3c --passes 1 > $O/r02j_prof_fps.log 2>&1
ls -la $O | grep r02j | head -30
