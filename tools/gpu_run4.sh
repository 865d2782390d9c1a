#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_row.json 2> $O/b_row.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --edge 640 > $O/b_640.json 2> $O/b_640.err; echo "bench640 rc=$?"
