#!/bin/bash
# Round 2, GPU call T: the state the round is handed in -- smoke, the whole GPU suite, the driver's two bench commands,
# the L / LC-L lines, the ncu launch list of the bench command and the ncu --set full capture of the LC conv launches.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02t_smoke.log 2>&1
echo "smoke exit $?" | tee $O/r02t_summary.txt
tail -n 1 $O/r02t_smoke.log | tee -a $O/r02t_summary.txt
timeout 900 python -m pytest tests -m gpu -q > $O/r02t_pytest_all.log 2>&1
echo "gpu suite exit $?" | tee -a $O/r02t_summary.txt
tail -n 3 $O/r02t_pytest_all.log | tee -a $O/r02t_summary.txt
cp $O/parity_abs_err.json $O/r02t_parity_abs_err.json 2>/dev/null
timeout 900 python bench.py > $O/r02t_bench_default.json 2>$O/r02t_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02t_bench_reference.json 2>$O/r02t_bench_reference.err
B="--no-cpu-baseline --no-cuda-baseline"
timeout 300 python bench.py --workload LC --steps 20 --warmup 5 $B --breakdown $O/r02t_breakdown_LC_S.json > $O/r02t_bench_LC_S.json 2>$O/r02t_bench_LC_S.err
timeout 300 python bench.py --workload LC --profile L --steps 10 --warmup 3 $B > $O/r02t_bench_LC_L.json 2>$O/r02t_bench_LC_L.err
timeout 300 python bench.py --workload L --steps 40 --warmup 10 $B > $O/r02t_bench_L_S.json 2>$O/r02t_bench_L_S.err
timeout 300 python bench.py --workload L --profile L --steps 20 --warmup 5 $B > $O/r02t_bench_L_L.json 2>$O/r02t_bench_L_L.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02t_launches_LC.csv \
  python bench.py --steps 2 --warmup 3 $B > $O/r02t_launches_LC.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_fwd_sb -s 37 -c 37 -f -o $O/r02t_prof_conv_LC_S \
  python tools/prof_conv.py --lc > $O/r02t_prof_conv_LC.log 2>&1
tail -2 $O/r02t_prof_conv_LC.log
python - <<'PY' | tee -a gpurun_out/r02t_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02t_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'], 3), d['unit'], round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 3),
              '; frac', r.get('frac'), '; kernel ms', r.get('kernel_ms_per_step'), '; cuda', (d.get('cuda_baseline') or {}).get('value'),
              '; cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
