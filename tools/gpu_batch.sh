#!/bin/bash
# slices per internal batch (IU_AUTO_BATCH): activation tensors of 74 slices exceed L2, smaller batches keep
# producer -> consumer traffic on chip but double the launches
for b in ${BATCHES:-74 111 128 148 256}; do
IU_AUTO_BATCH=$b timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch', $b, round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1), 'conv', round(d['kernel_ms_per_step']['conv'],2), 'stem', round(d['kernel_ms_per_step']['stem'],2), 'pool', round(d['kernel_ms_per_step']['pool'],2), 'clk', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])"
done
