#!/bin/bash
# round-2 evidence: smoke, ncu launch list of the bench command, --set full captures of the main kernels, GEMM DRAM traffic
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-decoder --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-1100} --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1; head -30 gpurun_out/r02_launches_summary.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_bf16_tc_kernel --clock-control none -c 330 --csv --log-file gpurun_out/r02_gemm_traffic.csv $CMD > gpurun_out/r02_ncu_traffic.log 2>&1; echo "traffic rc=$?"
python scripts/gemm_traffic.py gpurun_out/r02_gemm_traffic.csv gpurun_out/r02_gemm_traffic.json 54 2>&1 | tail -3
for spec in "gemm_bf16_tc_kernel:60:gemm:4" "agg_fwd_mma_kernel:3:aggfwd:1" "agg_bwd_img_kernel:3:aggbwdimg:1" "softmax_bwd_mma_kernel:3:smbwd:1" "softmax_fwd_mma_kernel:3:smfwd:1"; do
  IFS=: read k skip name cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -f -o gpurun_out/r02_prof_$name $CMD > gpurun_out/r02_ncu_full_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i gpurun_out/r02_prof_$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$name.csv 2>/dev/null
  rm -f gpurun_out/r02_prof_$name.ncu-rep      # (gpurun brings back at most 64 MiB: keep the CSV pages only)
done
ls -la gpurun_out | grep r02_ | head -40
