#!/bin/bash
# ncu --set full of one warm network pass (43 launches: stem, pool, 40 convs, fused tail) -> per-kernel table + DRAM bytes
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
ncu --set full --clock-control none -k regex:'conv_|maxpool' -s 43 -c 43 -o $O/r02_full_pass \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full_pass.log 2>&1; echo "ncu pass rc=$?"
ncu -i $O/r02_full_pass.ncu-rep --page raw --csv > $O/r02_full_pass_raw.csv 2>/dev/null; rm -f $O/r02_full_pass.ncu-rep
python tools/ncu_table.py $O/r02_full_pass_raw.csv > $O/r02_ncu_full_conv_pass.txt 2>&1; tail -50 $O/r02_ncu_full_conv_pass.txt
