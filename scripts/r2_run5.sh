#!/bin/bash
# full GPU suite (all failures listed) + bench A/B lines; outputs tagged by $TAG.  VARIANTS: "name:ENV=VAL ..." entries.
mkdir -p gpurun_out
TAG=${TAG:-r2w}
BENCH="python bench.py --steps 20 --warmup 5 --no-decoder --no-cpu-baseline --no-gpu-baseline"
if [ -z "${SKIP_TESTS}" ]; then
timeout 900 python -m pytest ${TESTS:-tests} -q -m gpu --timeout 600 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/gputests_${TAG}.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/gputests_${TAG}.log | tail -25
fi
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'], 4), round(d['e2e']['value']), d['roofline']['frac'], d.get('kernel_time_share_pct'))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
timeout 300 $BENCH > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_${TAG}.err; summ gpurun_out/bench_${TAG}.json
for v in ${VARIANTS}; do
  name=${v%%:*}; envs=${v#*:}
  timeout 300 env ${envs//,/ } $BENCH > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err; echo "bench $name rc=$?"
  summ gpurun_out/bench_${TAG}_${name}.json
done
