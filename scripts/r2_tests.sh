#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider -s ${PYTEST_ARGS} > gpurun_out/gputests_r2c.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|Error|error|assert" gpurun_out/gputests_r2c.log | tail -15
