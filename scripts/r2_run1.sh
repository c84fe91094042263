#!/bin/bash
# round-2 GPU call 1: GEMM probe (ablations, timelines, mixed formats), bench with in-graph timing + eager comparator, GPU tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python scripts/gemm_probe.py gpurun_out/r02_gemm_probe.json > gpurun_out/gemm_probe.log 2>&1; echo "probe rc=$?"
tail -5 gpurun_out/gemm_probe.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider > gpurun_out/gputests_r2a.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gputests_r2a.log
