#!/bin/bash
# Round 2, GPU call L (2 GPUs): BASELINE configs[4] train step and the LC forward at N=1 and N=2 on the same box.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
for prec in bf16x3c bf16; do
  timeout 600 python bench.py --workload train --precision $prec --steps 10 --warmup 3 > $O/r02l_bench_train_${prec}_1gpu.json 2>$O/r02l_bench_train_${prec}_1gpu.err
  timeout 600 $TR bench.py --gpus 2 --workload train --precision $prec --steps 10 --warmup 3 > $O/r02l_bench_train_${prec}_2gpu.json 2>$O/r02l_bench_train_${prec}_2gpu.err
done
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02l_bench_LC_2gpu.json 2>$O/r02l_bench_LC_2gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cuda-baseline > $O/r02l_bench_LC_1gpu.json 2>$O/r02l_bench_LC_1gpu.err
timeout 300 python -m pytest tests -m gpu -q -k "train or exchange or dist" > $O/r02l_pytest_train.log 2>&1
python - <<'PY' | tee gpurun_out/r02l_summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r02l_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['n_gpus'], 'gpu', round(d['value'], 2), d['unit'], round(d['ms_per_step'], 2), 'ms', 'exchange', d.get('exchange'))
    except Exception as e:
        print(f, 'unparsed', e)
PY
tail -3 $O/r02l_pytest_train.log
tail -3 $O/r02l_bench_train_bf16x3c_2gpu.err
