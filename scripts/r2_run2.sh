#!/bin/bash
# bench (in-graph timing) + GPU tests; outputs tagged by $TAG
mkdir -p gpurun_out
TAG=${TAG:-r2d}
timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}.err
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/gputests_${TAG}.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gputests_${TAG}.log
