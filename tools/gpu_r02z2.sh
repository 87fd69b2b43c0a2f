#!/bin/bash
# Round 2, GPU call Z2: compute-sanitizer synccheck + initcheck on the split-operand convolution tests
set -u
mkdir -p gpurun_out
O=gpurun_out
K="split_operand_conv_matches_oracle or spconv1x_golden"
timeout 200 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_zz_train_gpu.py tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/r02z_synccheck_conv.log 2>&1
echo "synccheck exit $?" | tee $O/r02z2_summary.txt
tail -n 3 $O/r02z_synccheck_conv.log | tee -a $O/r02z2_summary.txt
timeout 200 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_zz_train_gpu.py tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/r02z_initcheck_conv.log 2>&1
echo "initcheck exit $?" | tee -a $O/r02z2_summary.txt
grep -c "Uninitialized" $O/r02z_initcheck_conv.log | tee -a $O/r02z2_summary.txt
tail -n 3 $O/r02z_initcheck_conv.log | tee -a $O/r02z2_summary.txt
grep -m 12 -A6 "Uninitialized" $O/r02z_initcheck_conv.log | head -60
