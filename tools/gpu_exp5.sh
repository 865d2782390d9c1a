#!/bin/bash
O=gpurun_out
for v in 0 1; do
  echo "=== IU_ROW_STREAM64=$v"
  IU_ROW_STREAM64=$v IU_CONV_DEBUG=1 timeout 200 python tools/profile_forward.py --batch 74 --iters 2 2>&1 | grep -E "^ +(0|1|2|36|37) " | cut -c1-165
  IU_ROW_STREAM64=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1), 'clk', d['clocks']['sm_mhz'])"
done
