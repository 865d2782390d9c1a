#!/bin/bash
# Final-code multi-GPU check at G GPUs: NCCL bit identity (tests/test_multi_gpu.py), then the 1024^3 bench line.
set -u
O=gpurun_out; mkdir -p $O
G=$1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/pytest_multi_${G}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_multi_${G}gpu.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $G --steps 5 --warmup 3 > $O/r02_bench_${G}gpu_1024.json 2> $O/r02_bench_${G}gpu_1024.err
echo "bench G=$G rc=$? $(python -c "import json;d=json.load(open('$O/r02_bench_${G}gpu_1024.json'));print(round(d['value']/1e9,3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],1), d['clocks']['sm_mhz'])")"
