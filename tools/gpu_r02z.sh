#!/bin/bash
# Round 2, GPU call Z: compute-sanitizer memcheck on the split-operand convolution (both schedules, hand-off slots), the
# hash-path modality split, and one LC forward through the native GMA stage.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_zz_train_gpu.py tests/test_gpu_parity.py -m gpu -q -x \
  -k "split_operand_conv_matches_oracle or modality_split_matches_reference_golden or spconv1x_golden" > $O/r02z_memcheck_conv.log 2>&1
echo "memcheck conv/split exit $?" | tee $O/r02z_summary.txt
tail -n 3 $O/r02z_memcheck_conv.log | tee -a $O/r02z_summary.txt
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "test_native_gma_stage_and_overlapped_schedule_equal_module_path and True" > $O/r02z_memcheck_lc.log 2>&1
echo "memcheck LC exit $?" | tee -a $O/r02z_summary.txt
tail -n 3 $O/r02z_memcheck_lc.log | tee -a $O/r02z_summary.txt
