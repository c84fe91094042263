#!/bin/bash
# selected GPU tests + one bench; outputs tagged by $TAG
mkdir -p gpurun_out
TAG=${TAG:-r2e}
timeout 900 python -m pytest ${TESTS:-tests} -q -m gpu ${XFLAG--x} --timeout 600 -p no:cacheprovider -s ${PYTEST_ARGS} > gpurun_out/gputests_${TAG}.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|Error|error|assert|split3 err|worst grads|^B=" gpurun_out/gputests_${TAG}.log | tail -40
if [ -n "${BENCH_ARGS}" ]; then
timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('kernel_time_share_pct'))
PY
fi
