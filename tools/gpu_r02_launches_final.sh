#!/bin/bash
# ncu launch list (gpu__time_duration, no clock control) of one warm 74-slice pass on the final code
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_final.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch.log 2>&1
python tools/launch_table.py $O/launches_r02_final.csv > $O/r02_launches_final.txt 2>&1; tail -50 $O/r02_launches_final.txt
