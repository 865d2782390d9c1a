#!/bin/bash
# A/B of kernel variants: each line = env + label
set -u
O=gpurun_out; mkdir -p $O
V=interactive-unet_b200/build/variants
run() { label=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ab_$label.json 2> $O/ab_$label.err; python - <<EOF
import json
try:
    d=json.load(open("$O/ab_$label.json")); print("$label", round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,1), "Mvox/s conv", round(d["kernel_ms_per_step"]["conv"],2), "clk", d["clocks"]["sm_mhz"])
except Exception as e: print("$label FAILED", e)
EOF
}
run base IU_X=0
run pair IU_CONV_PAIR=1
run pairB16 IU_CONV_PAIR=1 IU_LIB=$V/libiunet_pairB16.so
run haloA2 IU_LIB=$V/libiunet_haloA2.so
run haloA2_all IU_CONV_VARIANT=2 IU_LIB=$V/libiunet_haloA2.so
run base2 IU_X=0
IU_CONV_PAIR=1 IU_LIB=$V/libiunet_pairB16.so IU_CONV_DEBUG=1 timeout 300 python tools/profile_forward.py --batch 74 --iters 2 > $O/dbg_pairB16.log 2>&1
IU_CONV_VARIANT=2 IU_LIB=$V/libiunet_haloA2.so IU_CONV_DEBUG=1 timeout 300 python tools/profile_forward.py --batch 74 --iters 2 > $O/dbg_haloA2.log 2>&1
