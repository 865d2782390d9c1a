#!/bin/bash
# ABBA timing of one environment switch on one B200: bash tools/gpu_ab.sh VAR A B [rounds]
# (run order A B B A ..., so that the box warming up under the power cap does not favour either side)
set -u
V=$1; A=$2; B=$3; N=${4:-2}
O=gpurun_out; mkdir -p $O
i=0
for r in $(seq 1 $N); do
  for m in $A $B $B $A; do
    i=$((i+1))
    env $V=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config0 > $O/ab_$i.json 2> $O/ab_$i.err
    python -c "import json;d=json.load(open('$O/ab_$i.json'));print('$V=$m', round(d['ms_per_step'],2),'ms  e2e', round(d['e2e']['ms_per_step'],2), ' conv', round(d['kernel_ms_per_step']['conv'],2), ' sm', d['clocks']['sm_mhz'], d['clocks'].get('power_w'))"
  done
done
