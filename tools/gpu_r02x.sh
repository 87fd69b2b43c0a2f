#!/bin/bash
# Round 2, GPU call X (8 GPUs): configs[4] train step and the LC forward at N = 8.
set -u
mkdir -p gpurun_out
O=gpurun_out
export MSMD_BENCH_HANG_DUMP=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
B="--no-cpu-baseline --no-cuda-baseline"
timeout 150 $TR bench.py --gpus 8 --workload train --precision bf16 --steps 10 --warmup 3 $B > $O/r02x_bench_train_bf16_8gpu.json 2>$O/r02x_bench_train_bf16_8gpu.err
echo "train bf16 N=8 exit $?" | tee $O/r02x_summary.txt
timeout 120 $TR bench.py --gpus 8 --steps 20 --warmup 5 $B > $O/r02x_bench_LC_S_8gpu.json 2>$O/r02x_bench_LC_S_8gpu.err
echo "LC N=8 exit $?" | tee -a $O/r02x_summary.txt
python - <<'PY' | tee -a gpurun_out/r02x_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02x_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['n_gpus'], 'gpu', round(d['value'], 2), d['unit'], round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 2), '; exchange', d.get('gradient_exchange'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
