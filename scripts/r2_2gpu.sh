#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -q -m gpu -x --timeout 600 -p no:cacheprovider -s > gpurun_out/dp_nccl.log 2>&1; echo "dp test rc=$?"
grep -E "passed|failed|Error|assert|graph \{|eager \{" gpurun_out/dp_nccl.log | tail -12
for v in base noearly f32wire; do
  unset EKAID_B200_EARLY_REDUCE EKAID_B200_AR_BF16
  if [ $v = noearly ]; then export EKAID_B200_EARLY_REDUCE=0; fi
  if [ $v = f32wire ]; then export EKAID_B200_AR_BF16=0; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_2gpu_$v.json 2> gpurun_out/bench_2gpu_$v.err; echo "bench2 $v rc=$?"
done
unset EKAID_B200_EARLY_REDUCE EKAID_B200_AR_BF16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/timeline.py > gpurun_out/timeline_2gpu.log 2>&1; echo "timeline rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"
