#!/bin/bash
# compute-sanitizer memcheck over the kernels added / rewritten in round 2 (small cases: the tool is 10-50x slower)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "$@" -q -m gpu -x -p no:cacheprovider > gpurun_out/r02_sanitizer_$tag.log 2>&1; echo "$tag rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_$tag.log | tail -2; }
run decoder tests/test_gpu_speaker.py -k "(teacher_forced and bf16-3) or stop_condition"
run skinny tests/test_gpu_kernels.py -k "(skinny and 17-304-72) or (skinny and 1-148-512) or semantic_labels or split3 and 520"
run edge tests/test_gpu_parity.py -k "test_gradients_match_oracle and c1_b3 and bf16"
run labels tests/test_gpu_step.py -k "int8_label and c4"
