#!/bin/bash
# ncu --set full of the warm layer1.0 conv1 / conv2 launches (row-folded kernel, TMA-filled A ring)
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
ncu --set full --import-source on --clock-control none -k regex:'conv_row' -s 10 -c 2 -o $O/r02_full_row_tma \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full_row.log 2>&1; echo "ncu rc=$?"; ls -la $O/r02_full_row_tma.ncu-rep
