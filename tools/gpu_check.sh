#!/bin/bash
# Development loop on a B200 (through gpurun): row-kernel conv tests, the whole parity suite, a short bench, role counters.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_conv and row" > $O/pytest_row.log 2>&1; echo "row conv rc=$?"; tail -4 $O/pytest_row.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_row.json 2> $O/b_row.err; echo "bench rc=$?"
IU_CONV_DEBUG=1 timeout 300 python tools/profile_forward.py --batch 74 --iters 2 > $O/dbg_row.log 2>&1; tail -16 $O/dbg_row.log | cut -c1-165
