#!/bin/bash
# ncu evidence for the bench command (GPU box): per-launch device times of one bench step and a full-section capture
# of the dominant kernels.  Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-60000} --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_run.log 2>&1
echo "launch list rc=$? rows=$(wc -l < gpurun_out/launches.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-mv_stage}" -s ${NCU_SKIP:-400} -c ${NCU_FULL_COUNT:-4} \
    -f -o gpurun_out/prof_matvec $CMD > gpurun_out/prof_run.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
