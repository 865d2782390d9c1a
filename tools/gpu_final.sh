#!/bin/bash
# End-of-round evidence: parity suite, smoke, default bench (with CPU baseline), reference arm, launch list, role counters.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
python bench.py --steps 5 --warmup 3 > $O/b_final.json 2> $O/b_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/b_final_ref.json 2> $O/b_final_ref.err; echo "ref rc=$?"
IU_CONV_DEBUG=1 timeout 300 python tools/profile_forward.py --batch 74 --iters 2 > $O/dbg_final.log 2>&1
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_final.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch.log 2>&1
timeout 600 python bench.py --edge 1024 --classes 4 --steps 2 --warmup 1 --no-cpu-baseline > $O/b_1024c4.json 2> $O/b_1024c4.err; echo "1024^3 C=4 rc=$?"
timeout 300 python bench.py --encoder resnet18 --steps 5 --warmup 3 --no-cpu-baseline > $O/b_resnet18.json 2> $O/b_resnet18.err; echo "resnet18 rc=$?"
timeout 300 python tools/zarr_bench.py > $O/zarr_bench.json 2> $O/zarr_bench.err; echo "zarr disk-to-disk rc=$?"
