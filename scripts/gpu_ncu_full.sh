#!/bin/bash
# ncu --set full captures of the main kernels (one launch each, steady state), eager launches
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph"
for spec in "gemm_bf16_tc_kernel:40:gemm" "agg_fwd_mma_kernel:3:aggfwd" "agg_bwd_img_kernel:3:aggbwdimg" "softmax_bwd_mma_kernel:3:smbwd" "gru_seq_bwd_kernel:1:grubwd"; do
  IFS=: read k skip name <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c ${CNT:-2} -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_full_$name.log 2>&1
  echo "$name rc=$?"
done
