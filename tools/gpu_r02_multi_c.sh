#!/bin/bash
# Refresh of the config-3 / config-5 lines at G GPUs with the pipelined end-to-end path:  bash tools/gpu_r02_multi_c.sh G "tags"
set -u
O=gpurun_out; mkdir -p $O
G=$1
run() {
  local tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $G "$@" > $O/r02_bench_${G}gpu_${tag}.json 2> $O/r02_bench_${G}gpu_${tag}.err
  echo "bench G=$G $tag rc=$? $(python -c "import json;d=json.load(open('$O/r02_bench_${G}gpu_${tag}.json'));print(round(d['value']/1e9,3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],1))")"
}
run 2048 --edge 2048 --steps 2 --warmup 1
if [ "$G" == "8" ]; then
  run 1024_c4 --edge 1024 --classes 4 --steps 5 --warmup 2
  run 2048_c4 --edge 2048 --classes 4 --steps 2 --warmup 1
else
  run 1024 --steps 10 --warmup 3
fi
