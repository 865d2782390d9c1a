#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; V=interactive-unet_b200/build/variants
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_nofence$i.json 2> $O/b_nofence.err; echo "bench rc=$?"
IU_LIB=$V/libiunet_fence.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_fence$i.json 2> $O/b_fence.err
done
python - <<EOF
import json
for f in ["b_nofence1","b_fence1","b_nofence2","b_fence2"]:
    d=json.load(open("$O/"+f+".json")); print(f, round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,1), "Mvox/s frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"])
EOF
IU_CONV_DEBUG=1 timeout 300 python tools/profile_forward.py --batch 74 --iters 2 > $O/dbg_row.log 2>&1; tail -16 $O/dbg_row.log | cut -c1-160
