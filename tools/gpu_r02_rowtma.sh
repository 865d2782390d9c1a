#!/bin/bash
# A/B of the TMA-filled A ring of the Cout-64 row layers (IU_ROW_TMA = 0 / 1 / 2) on one B200.
set -u
O=gpurun_out; mkdir -p $O
for m in 1 2; do
  IU_ROW_TMA=$m timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_conv_matches and row_k64n64 and not pertap and not halo and not pair" > $O/rowtma_conv_$m.log 2>&1
  echo "IU_ROW_TMA=$m conv rc=$?"; tail -2 $O/rowtma_conv_$m.log | cut -c1-200
done
good=0
grep -q " passed" $O/rowtma_conv_1.log && ! grep -q failed $O/rowtma_conv_1.log && good=1
[ $good = 0 ] && grep -q " passed" $O/rowtma_conv_2.log && ! grep -q failed $O/rowtma_conv_2.log && good=2
echo "working mode: $good"
[ $good = 0 ] && exit 0
export IU_ROW_TMA=$good
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_matches or resnet18 or predict_block or voxel or fusion" > $O/rowtma_fwd.log 2>&1; echo "forward rc=$?"; tail -2 $O/rowtma_fwd.log | cut -c1-200
for m in 0 $good; do
  IU_ROW_TMA=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config0 > $O/b_rowtma_$m.json 2> $O/b_rowtma_$m.err
  python -c "import json;d=json.load(open('$O/b_rowtma_$m.json'));print('IU_ROW_TMA=$m', round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['ms_per_step'],2), 'conv', round(d['kernel_ms_per_step']['conv'],2), d['clocks']['sm_mhz'])"
done
for r in 1 2; do
  IU_ROW_RES_TMA=$r python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config0 > $O/b_rowres_$r.json 2> $O/b_rowres_$r.err
  python -c "import json;d=json.load(open('$O/b_rowres_$r.json'));print('IU_ROW_RES_TMA=$r', round(d['ms_per_step'],2),'ms conv', round(d['kernel_ms_per_step']['conv'],2), d['clocks']['sm_mhz'])"
done
IU_CONV_DEBUG=1 python tools/profile_forward.py --batch 74 > $O/role_rowtma.txt 2>&1; sed -n 4,12p $O/role_rowtma.txt
