#!/bin/bash
# per-source-line stall samples of the edge kernels (ncu --set full --import-source on -> --page source --csv)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-decoder --no-graph"
for spec in ${SPECS:-"softmax_fwd_mma_kernel:3:smfwd" "softmax_bwd_mma_kernel:3:smbwd" "agg_fwd_mma_kernel:3:aggfwd" "agg_bwd_img_kernel:3:aggbwdimg"}; do
  IFS=: read k skip name <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02_src_$name $CMD > gpurun_out/r02_ncu_src_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i gpurun_out/r02_src_$name.ncu-rep --page source --csv > gpurun_out/r02_ncu_source_$name.csv 2>gpurun_out/r02_ncu_source_$name.err
  rm -f gpurun_out/r02_src_$name.ncu-rep
  ls -la gpurun_out/r02_ncu_source_$name.csv
done
