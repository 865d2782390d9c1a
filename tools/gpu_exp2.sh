#!/bin/bash
O=gpurun_out; V=interactive-unet_b200/build/variants
for v in backoff32 backoff200; do
  echo "=== $v"
  IU_LIB=$V/libiunet_$v.so IU_CONV_DEBUG=1 timeout 200 python tools/profile_forward.py --batch 74 --iters 2 2>&1 | grep -E "^ +(0|1|16|34|38|39|40|41|42) " | cut -c1-150
  IU_LIB=$V/libiunet_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2),'ms', d['clocks']['sm_mhz'])"
done
