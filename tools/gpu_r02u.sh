#!/bin/bash
# Round 2, GPU call U (2 GPUs): BASELINE configs[4] train step at N = 1 and N = 2 (gradient all-reduce over NVLink), LC and L
# forward at N = 2.  Every run under its own short timeout: a hang costs minutes on two GPUs.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
B="--no-cpu-baseline --no-cuda-baseline"
timeout 240 $TR bench.py --gpus 2 --workload train --precision bf16 --steps 10 --warmup 3 $B > $O/r02u_bench_train_bf16_2gpu.json 2>$O/r02u_bench_train_bf16_2gpu.err
echo "train bf16 N=2 exit $?" | tee $O/r02u_summary.txt
timeout 240 python bench.py --workload train --precision bf16 --steps 10 --warmup 3 $B > $O/r02u_bench_train_bf16_1gpu.json 2>$O/r02u_bench_train_bf16_1gpu.err
timeout 240 $TR bench.py --gpus 2 --workload train --steps 10 --warmup 3 $B > $O/r02u_bench_train_bf16x3c_2gpu.json 2>$O/r02u_bench_train_bf16x3c_2gpu.err
echo "train bf16x3c N=2 exit $?" | tee -a $O/r02u_summary.txt
timeout 240 python bench.py --workload train --steps 10 --warmup 3 $B > $O/r02u_bench_train_bf16x3c_1gpu.json 2>$O/r02u_bench_train_bf16x3c_1gpu.err
timeout 240 $TR bench.py --gpus 2 --steps 20 --warmup 5 $B > $O/r02u_bench_LC_S_2gpu.json 2>$O/r02u_bench_LC_S_2gpu.err
timeout 240 $TR bench.py --gpus 2 --workload L --steps 40 --warmup 10 $B > $O/r02u_bench_L_S_2gpu.json 2>$O/r02u_bench_L_S_2gpu.err
python - <<'PY' | tee -a gpurun_out/r02u_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02u_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['n_gpus'], 'gpu', round(d['value'], 2), d['unit'], round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 2), '; exchange', d.get('exchange'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
tail -3 $O/r02u_bench_train_bf16_2gpu.err
