#!/bin/bash
mkdir -p gpurun_out
N=${NG:-8}
run() { # tag, extra env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline ${BARGS} > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; echo "bench$N $tag rc=$?"
}
run nvls NCCL_ALGO=NVLS
run simple NCCL_PROTO=Simple
run tree NCCL_ALGO=Tree
run nvlstree NCCL_ALGO=NVLSTree
