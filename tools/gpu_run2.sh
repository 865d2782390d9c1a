#!/bin/bash
# row-kernel bring-up: conv unit tests (row variant), then the whole parity suite, then bench A/B
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_conv and row" > $O/pytest_row.log 2>&1; echo "row conv rc=$?"; tail -15 $O/pytest_row.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_row.json 2> $O/b_row.err; echo "bench rc=$?"
IU_CONV_ROW=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_norow.json 2> $O/b_norow.err
IU_CONV_DEBUG=1 timeout 300 python tools/profile_forward.py --batch 74 --iters 2 > $O/dbg_row.log 2>&1
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_row.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch.log 2>&1
