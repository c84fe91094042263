#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 300 -p no:cacheprovider > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/kernels.log
timeout 600 python scripts/gemm_probe.py gpurun_out/r02_gemm_probe_b.json > gpurun_out/gemm_probe.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/gemm_probe.log
