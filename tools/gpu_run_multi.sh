#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): the sharded bench at N ranks + a bit-identity check against 1 GPU
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/gpus_$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/check_sharded.py > $O/check_sharded_$N.log 2>&1; echo "check rc=$?"; tail -3 $O/check_sharded_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 > $O/b_gpus$N.json 2> $O/b_gpus$N.err; echo "bench rc=$?"; tail -c 600 $O/b_gpus$N.json
