#!/bin/bash
# Round-2 development loop on one B200: parity suite, smoke, short bench, latency.
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $O/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-config0 > $O/b_check.json 2> $O/b_check.err; echo "bench rc=$?"; tail -c 1500 $O/b_check.json
timeout 300 python tools/latency.py --iters 300 > $O/latency.json 2> $O/latency.err; echo "latency rc=$?"; cat $O/latency.json
