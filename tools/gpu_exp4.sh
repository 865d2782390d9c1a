#!/bin/bash
O=gpurun_out; V=interactive-unet_b200/build/variants
for v in base rot A2 rotA2; do
  echo "=== $v (halo kernel forced on the wide layers)"
  L=$V/libiunet_$v.so; [ $v = base ] && L=interactive-unet_b200/libiunet_b200.so
  IU_LIB=$L IU_CONV_VARIANT=2 IU_CONV_DEBUG=1 timeout 200 python tools/profile_forward.py --batch 74 --iters 2 2>&1 | grep -E "^ +(8|16|28|32|34|36) " | cut -c1-150
  IU_LIB=$L IU_CONV_VARIANT=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('halo-all', round(d['ms_per_step'],2),'ms conv', round(d['kernel_ms_per_step']['conv'],2), 'clk', d['clocks']['sm_mhz'])"
  IU_LIB=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('default ', round(d['ms_per_step'],2),'ms conv', round(d['kernel_ms_per_step']['conv'],2), 'clk', d['clocks']['sm_mhz'])"
done
