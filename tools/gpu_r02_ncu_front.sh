#!/bin/bash
# ncu --set full (with source counters) of the warm stem and max-pool launches of one 74-slice pass
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
ncu --set full --import-source on --clock-control none -k regex:'conv_stem|maxpool' -s 2 -c 2 -o $O/r02_full_front \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full_front.log 2>&1; echo "ncu rc=$?"; ls -la $O/r02_full_front.ncu-rep
