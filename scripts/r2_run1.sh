#!/bin/bash
# bench with in-graph timing + GPU tests
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2b.err
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider > gpurun_out/gputests_r2b.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gputests_r2b.log
