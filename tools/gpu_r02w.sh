#!/bin/bash
# Round 2, GPU call W (2 GPUs): configs[4] train step at N = 2 after the fix of bench.py's rank-0 profile pass.
set -u
mkdir -p gpurun_out
O=gpurun_out
export MSMD_BENCH_HANG_DUMP=120
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
B="--no-cpu-baseline --no-cuda-baseline"
timeout 170 $TR bench.py --gpus 2 --workload train --precision bf16 --steps 10 --warmup 3 $B > $O/r02w_bench_train_bf16_2gpu.json 2>$O/r02w_bench_train_bf16_2gpu.err
echo "train bf16 N=2 exit $?" | tee $O/r02w_summary.txt
timeout 170 $TR bench.py --gpus 2 --workload train --steps 10 --warmup 3 $B > $O/r02w_bench_train_bf16x3c_2gpu.json 2>$O/r02w_bench_train_bf16x3c_2gpu.err
echo "train bf16x3c N=2 exit $?" | tee -a $O/r02w_summary.txt
python - <<'PY' | tee -a gpurun_out/r02w_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02w_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['n_gpus'], 'gpu', round(d['value'], 2), d['unit'], round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value'], 2), '; exchange', d.get('gradient_exchange'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
grep -A25 "Thread 0x\|most recent call first" $O/r02w_bench_train_bf16_2gpu.err | head -80
