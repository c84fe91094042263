#!/bin/bash
# final round-2 pass: full GPU suite, bench A/B lines, ncu launch list of the bench command, --set full captures of the
# edge kernels that changed after the r02_* captures (one ncu run: 9 consecutive launches cover all three kernels)
mkdir -p gpurun_out
TAG=${TAG:-r2y}
TAG=$TAG VARIANTS="${VARIANTS}" bash scripts/r2_run5.sh
if [ -n "${SKIP_NCU}" ]; then exit 0; fi
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-decoder --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-1100} --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/r02b_launches.csv > gpurun_out/r02b_launches_summary.txt 2>&1; head -12 gpurun_out/r02b_launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:agg_bwd_img_kernel|agg_fwd_mma_kernel|softmax_bwd_mma_kernel' -s 9 -c 9 -f -o gpurun_out/r02b_prof_edge $CMD > gpurun_out/r02b_ncu_full_edge.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02b_prof_edge.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_full_edge.csv 2>/dev/null
ncu -i gpurun_out/r02b_prof_edge.ncu-rep --page source --csv -k regex:agg_bwd_img_kernel > gpurun_out/r02b_ncu_source_aggbwdimg.csv 2>/dev/null
rm -f gpurun_out/r02b_prof_edge.ncu-rep
ls -la gpurun_out | grep r02b_
