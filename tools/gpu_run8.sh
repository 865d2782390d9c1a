#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
b() { label=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > $O/ab_$label.json; python -c "
import json; d=json.load(open('$O/ab_$label.json')); print('$label', round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1), 'conv', round(d['kernel_ms_per_step']['conv'],2), 'clk', d['clocks']['sm_mhz'])"; }
b bn256 IU_X=1
b bn128 IU_CONV_BN256=0
b bn256b IU_X=1
b bn128b IU_CONV_BN256=0
