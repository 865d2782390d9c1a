#!/bin/bash
# Round-2 evidence on one B200: final bench (+ reference arm), launch list, ncu --set full of one network pass and of the
# HBM-bound kernels, one source-level capture of the fused decoder tail.  Only CSV exports and one .ncu-rep come back.
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
python bench.py --steps 10 --warmup 3 > $O/r02_bench_final.json 2> $O/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref rc=$?"
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_final.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch.log 2>&1
python tools/launch_table.py $O/launches_r02_final.csv > $O/r02_launches_final.txt 2>&1; tail -3 $O/r02_launches_final.txt
# one warm network pass: 43 launches of the repository's kernels (stem, pool, 40 convs, fused tail)
ncu --set full --clock-control none -k regex:'iu::' -s 43 -c 43 -o $O/r02_full_pass \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full_pass.log 2>&1; echo "ncu pass rc=$?"
ncu -i $O/r02_full_pass.ncu-rep --page raw --csv > $O/r02_full_pass_raw.csv 2>/dev/null; rm -f $O/r02_full_pass.ncu-rep
python tools/ncu_table.py $O/r02_full_pass_raw.csv > $O/r02_ncu_full_conv_pass.txt 2>&1; tail -4 $O/r02_ncu_full_conv_pass.txt
# HBM-bound kernels of a 256^3 volume (second prediction = warm): gather x3 orientations, reduce, pool, stem
python tools/profile_volume.py > $O/pv_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'gather_|reduce_|maxpool|conv_stem' -s 10 -c 10 -o $O/r02_full_hbm \
    python tools/profile_volume.py > $O/ncu_full_hbm.log 2>&1; echo "ncu hbm rc=$?"
ncu -i $O/r02_full_hbm.ncu-rep --page raw --csv > $O/r02_full_hbm_raw.csv 2>/dev/null; rm -f $O/r02_full_hbm.ncu-rep
python tools/ncu_table.py $O/r02_full_hbm_raw.csv > $O/r02_ncu_full_hbm_kernels.txt 2>&1; cat $O/r02_ncu_full_hbm_kernels.txt
# the fused decoder tail with source correlation
ncu --set full --clock-control none --import-source on -k regex:'conv_chain' -s 1 -c 1 -o $O/r02_full_chain \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full_chain.log 2>&1; echo "ncu chain rc=$?"
ncu -i $O/r02_full_chain.ncu-rep --page raw --csv > $O/r02_full_chain_raw.csv 2>/dev/null
ls -la $O/*.ncu-rep
