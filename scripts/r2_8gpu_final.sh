#!/bin/bash
# round-2 multi-GPU records: default batch 64, BASELINE config 3 (batch 256 per GPU, train), config 2 (batch 512 per GPU, inference)
mkdir -p gpurun_out
N=${NG:-8}
run() { # tag, bench args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-decoder "$@" > gpurun_out/r02_bench_${N}gpu_$tag.json 2> gpurun_out/r02_bench_${N}gpu_$tag.err; echo "bench$N $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_${N}gpu_$tag.json").read().strip().splitlines()[-1]); print("$tag", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
except Exception as e: print("$tag", "ERR", e)
PY
}
run b64
run b256 --batch 256
run infer_b512 --batch 512 --mode infer
