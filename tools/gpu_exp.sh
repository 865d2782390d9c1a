#!/bin/bash
O=gpurun_out; V=interactive-unet_b200/build/variants
for v in A_ALIGNED NO_GATHER NO_EPI; do
  echo "=== $v"
  IU_LIB=$V/libiunet_exp_$v.so IU_CONV_DEBUG=1 timeout 200 python tools/profile_forward.py --batch 74 --iters 2 2>&1 | grep -E "^ +(0|1|38|39|40|41|42) " | cut -c1-150
done
