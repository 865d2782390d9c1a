#!/bin/bash
# End-of-round-2 evidence with the final code on one B200: whole GPU suite, smoke(), the bench as the driver runs it
# (20 steps, 5 warm-ups) + reference arm, resnet18 line, 1024^3 C=4 on one GPU, store disk-to-disk stages.
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref rc=$?"
timeout 300 python bench.py --encoder resnet18 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/r02_bench_resnet18.json 2> $O/r02_bench_resnet18.err; echo "resnet18 rc=$?"
timeout 600 python bench.py --edge 1024 --classes 4 --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > $O/r02_bench_1gpu_1024_c4.json 2> $O/r02_bench_1gpu_1024_c4.err; echo "1024^3 C=4 rc=$?"
timeout 300 python tools/zarr_bench.py > $O/zarr_bench.out 2> $O/zarr_bench.err; echo "zarr rc=$?"; grep '^{' $O/zarr_bench.out | tail -1 > $O/r02_zarr_disk_to_disk_512.json  # the reference-style progress lines go to stdout too
for f in r02_bench_final r02_bench_resnet18 r02_bench_1gpu_1024_c4; do python -c "import json;d=json.load(open('$O/$f.json'));print('$f', round(d['value']/1e6,1),'Mvox/s', round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"; done
