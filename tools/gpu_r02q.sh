#!/bin/bash
# Round 2, GPU call Q: persistent kernel v3 (epilogue units in the work shares, sleeping waits) -- timeline + A/B
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/tc_trace.py --precision bf16x3c --sb-variant 0 --dump-cta --only "128->128 k27,32->32 k27,128->128 k3" --json $O/r02q_tc_trace_S_sbp.json > $O/r02q_tc_trace_S_sbp.txt 2>&1
timeout 600 python tools/sb_bench.py --lc --json $O/r02q_sb_bench_LC.json > $O/r02q_sb_bench_LC.txt 2>&1
tail -n 2 $O/r02q_sb_bench_LC.txt
MSMD_TC_TUNE="epi=1" timeout 600 python tools/sb_bench.py --lc --json $O/r02q_sb_bench_LC_epi0.json > $O/r02q_sb_bench_LC_epi0.txt 2>&1
tail -n 2 $O/r02q_sb_bench_LC_epi0.txt
MSMD_TC_TUNE="epi=17" timeout 600 python tools/sb_bench.py --lc --json $O/r02q_sb_bench_LC_epi16.json > $O/r02q_sb_bench_LC_epi16.txt 2>&1
tail -n 2 $O/r02q_sb_bench_LC_epi16.txt
