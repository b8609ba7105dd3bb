#!/bin/bash
# ncu evidence for the bench command (GPU box): per-launch device times of one bench step and full-section captures
# of the dominant kernels.  Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-60000} --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_run.log 2>&1
echo "launch list rc=$? rows=$(wc -l < gpurun_out/launches.csv)"
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -n 30 gpurun_out/launches_summary.txt
for spec in ${NCU_SPECS:-pcg_persistent_kernel:40:1 cholqr_kernel:40:1 stack_persistent_kernel:60:1}; do
  IFS=: read -r kern skip count <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s $skip -c $count \
      -f -o gpurun_out/prof_$kern $CMD > gpurun_out/prof_${kern}_run.log 2>&1
  echo "full capture $kern rc=$?"
  ncu -i gpurun_out/prof_$kern.ncu-rep --page raw --csv > gpurun_out/prof_${kern}_raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
