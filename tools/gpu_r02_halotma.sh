#!/bin/bash
# Parity, then ABBA timing and role counters, of the halo kernel with TMA-filled tiles (IU_HALO_TMA) on one B200.
# IU_HALO_TMA=1: 18-pixel halo rows; 24: rows padded to 24 pixels (every 8-row group starts a swizzle period).
set -u
O=gpurun_out; mkdir -p $O
good=0
for m in 1 24; do
  IU_HALO_TMA_TEST=$m timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_conv_matches and halo_tma" > $O/halotma_conv_$m.log 2>&1
  echo "pitch mode $m conv rc=$?"; tail -2 $O/halotma_conv_$m.log | cut -c1-200
  if grep -q " passed" $O/halotma_conv_$m.log && ! grep -q failed $O/halotma_conv_$m.log; then good=$m; break; fi
done
echo "working mode: $good"
[ $good = 0 ] && exit 0
IU_HALO_TMA=$good timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_matches or resnet18 or predict_block or voxel or batch_size" > $O/halotma_fwd.log 2>&1; echo "forward rc=$?"; tail -2 $O/halotma_fwd.log | cut -c1-200
grep -q failed $O/halotma_fwd.log && exit 0
bash tools/gpu_ab.sh IU_HALO_TMA 0 $good 1
IU_HALO_TMA=$good IU_CONV_DEBUG=1 python tools/profile_forward.py --batch 74 > $O/role_halotma.txt 2>&1; sed -n 4,45p $O/role_halotma.txt
