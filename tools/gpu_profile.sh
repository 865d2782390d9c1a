#!/bin/bash
# Evidence pass: default bench (with the CPU baseline), reference arm, launch list, full ncu captures of four kernels.
set -u
O=gpurun_out; mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/b_final.json 2> $O/b_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/b_final_ref.json 2> $O/b_final_ref.err; echo "ref rc=$?"
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_final.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch.log 2>&1
for spec in "layer1_row64:43" "layer3_pertap:59" "dec2conv1_rowstream:79" "dec4conv2_row16:84"; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:'conv_(row|tc|halo)' -s $skip -c 1 \
      -o $O/r01_full_$name python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full_$name.log 2>&1
  ncu -i $O/r01_full_$name.ncu-rep --page raw --csv > $O/r01_full_${name}_raw.csv 2>/dev/null
  ls -la $O/r01_full_$name.ncu-rep
done
