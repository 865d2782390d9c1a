#!/bin/bash
# odd volume edges (the weak-scaling sizes of bench.py): parity suite + single-GPU bench at 640 / 800 next to 512
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
for e in 512 640 800; do
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --edge $e 2>/dev/null > $O/b_edge$e.json; python -c "
import json; d=json.load(open('$O/b_edge$e.json')); print('edge', $e, round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1), 'Mvox/s frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"
done
