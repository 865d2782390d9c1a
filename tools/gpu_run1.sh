#!/bin/bash
# Round-1 GPU pass: parity suite, bench (both arms), launch list, one full ncu capture of the conv kernels.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py --steps 5 --warmup 3 > $O/b_default.json 2> $O/b_default.err; echo "bench rc=$?"
IU_CONV_PAIR=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_pair.json 2> $O/b_pair.err; echo "bench pair rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_default2.json 2> $O/b_default2.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/b_ref.json 2> $O/b_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_(halo|tc|pair)' -s 43 -c 43 \
    -o $O/r01_conv_full python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_full.log 2>&1
ls -la $O/*.ncu-rep
ncu -i $O/r01_conv_full.ncu-rep --page raw --csv > $O/r01_conv_full_raw.csv 2>/dev/null
SZ=$(stat -c %s $O/r01_conv_full.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 45000000 ]; then echo "rep too large ($SZ), keeping csv only"; rm -f $O/r01_conv_full.ncu-rep; fi
