#!/bin/bash
# Run on the GPU box (via gpurun): kernel tests first (isolated, bounded), then the parity suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 300 -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/kernels.log
tail -40 gpurun_out/kernels.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_mode.py tests/test_gpu_step.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/parity.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/parity.log | tail -60
