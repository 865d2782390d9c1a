#!/bin/bash
O=gpurun_out; V=interactive-unet_b200/build/variants
b() { label=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['ms_per_step'],2),'ms conv', round(d['kernel_ms_per_step']['conv'],2), 'clk', d['clocks']['sm_mhz'])"; }
b backoff32 IU_LIB=$V/libiunet_backoff32.so
b backoff32_halo_all IU_LIB=$V/libiunet_backoff32.so IU_CONV_VARIANT=2
b backoff32_pair IU_LIB=$V/libiunet_backoff32.so IU_CONV_PAIR=1
b backoff32_again IU_LIB=$V/libiunet_backoff32.so
IU_LIB=$V/libiunet_backoff32.so IU_CONV_VARIANT=2 IU_CONV_DEBUG=1 timeout 200 python tools/profile_forward.py --batch 74 --iters 2 2>&1 | grep -E "^ +(8|9|16|17|28|29|32|33|34|35|36) " | cut -c1-150
echo pair
IU_LIB=$V/libiunet_backoff32.so IU_CONV_PAIR=1 IU_CONV_DEBUG=1 timeout 200 python tools/profile_forward.py --batch 74 --iters 2 2>&1 | grep -E "^ +(8|9|16|17|28|29|32|33|34|35|36) " | cut -c1-150
