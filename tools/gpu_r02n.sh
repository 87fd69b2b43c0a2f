#!/bin/bash
# Round 2, GPU call N: per-role timeline of the persistent split-operand kernel (why is it not faster than tile-per-CTA?)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/tc_trace.py --precision bf16x3c --sb-variant 0 --dump-cta --only "128->128 k27,32->32 k27,128->128 k3,64->64 k27" --json $O/r02n_tc_trace_S_sbp.json > $O/r02n_tc_trace_S_sbp.txt 2>&1
tail -n 5 $O/r02n_tc_trace_S_sbp.txt
timeout 300 python tools/tc_trace.py --precision bf16x3c --sb-variant 1 --json $O/r02n_tc_trace_S_sb_tile.json > $O/r02n_tc_trace_S_sb_tile.txt 2>&1
tail -n 15 $O/r02n_tc_trace_S_sb_tile.txt
