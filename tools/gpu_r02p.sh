#!/bin/bash
# Round 2, GPU call P: ncu --set full with source correlation of the persistent kernel on the 128->128 SubM layers
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spconv_fwd_sbp --launch-skip 38 -c 2 \
  -o $O/r02p_sbp_128 -f python tools/prof_conv.py > $O/r02p_prof.log 2>&1
tail -n 5 $O/r02p_prof.log
ls -la $O/r02p_sbp_128.ncu-rep
