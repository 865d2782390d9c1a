#!/bin/bash
# Launch list of one network pass (74 slices of 512^2) + role of the new kernels; A/B of the CTA-pair per-tap kernel.
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch.log 2>&1
python tools/launch_table.py $O/launches_r02.csv > $O/launches_r02.txt 2>&1; cat $O/launches_r02.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pertap_pair2 or fp16_range" > $O/pytest_pair2.log 2>&1; echo "pair2 tests rc=$?"; tail -5 $O/pytest_pair2.log
for v in 0 1 2 3; do
  IU_CONV_PAIR2=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/b_pair2_$v.json 2> $O/b_pair2_$v.err
  echo "IU_CONV_PAIR2=$v rc=$? $(python -c "import json;d=json.load(open('$O/b_pair2_$v.json'));print(round(d['ms_per_step'],2),'ms conv',round(d['roofline']['kernel_ms_per_step'],2),'clk',d['clocks']['sm_mhz'])")"
done
IU_CONV_PAIR2=3 python tools/profile_forward.py --batch 74 --iters 2 > $O/pf_pair2.log 2>&1 &&
IU_CONV_PAIR2=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_pair2.csv \
    python tools/profile_forward.py --batch 74 --iters 2 > $O/ncu_launch2.log 2>&1
python tools/launch_table.py $O/launches_r02_pair2.csv $O/launches_r02.csv > $O/launches_r02_pair2.txt 2>&1; cat $O/launches_r02_pair2.txt
