#!/bin/bash
# Round-2 multi-GPU evidence on one box with >= $1 GPUs:  bash tools/gpu_r02_multi.sh "2" | "4 8"
# per world size G: NCCL bit-identity test, default bench (1024^3, C=2), config 5 (2048^3), and at G=8 config 3 (1024^3, C=4)
set -u
O=gpurun_out; mkdir -p $O
run() {  # G, tag, bench args...
  local G=$1 tag=$2; shift; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $G "$@" > $O/r02_bench_${G}gpu_${tag}.json 2> $O/r02_bench_${G}gpu_${tag}.err
  echo "bench G=$G $tag rc=$? $(head -c 400 $O/r02_bench_${G}gpu_${tag}.json)"
  tail -3 $O/r02_bench_${G}gpu_${tag}.err | cut -c1-300
}
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
for G in $1; do
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "[$G]" > $O/pytest_multi_${G}.log 2>&1; echo "pytest world=$G rc=$?"; tail -3 $O/pytest_multi_${G}.log
  run $G 1024 --steps 10 --warmup 3
  run $G 2048 --edge 2048 --steps 2 --warmup 1
  if [ "$G" == "8" ]; then
    run $G 1024_c4 --edge 1024 --classes 4 --steps 5 --warmup 2
    run $G 2048_c4 --edge 2048 --classes 4 --steps 2 --warmup 1
  fi
done
