#!/bin/bash
# Re-check of the sharded path after the pipelined host tail: NCCL bit-identity test + default bench at G GPUs.
set -u
O=gpurun_out; mkdir -p $O
G=$1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "[$G]" > $O/pytest_multi_${G}.log 2>&1; echo "pytest world=$G rc=$?"; tail -3 $O/pytest_multi_${G}.log
grep -c "pipelined tail): bit-identical on every rank = True" $O/sharded_check_${G}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $G --steps 10 --warmup 3 > $O/r02_bench_${G}gpu_1024_b.json 2> $O/r02_bench_${G}gpu_1024_b.err
echo "bench rc=$? $(python -c "import json;d=json.load(open('$O/r02_bench_${G}gpu_1024_b.json'));print(round(d['value']/1e9,3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],1))")"
tail -2 $O/r02_bench_${G}gpu_1024_b.err | cut -c1-300
