#!/bin/bash
# ncu evidence for the bench command (GPU box): per-launch device times of one C3 bench step, full-section captures of the
# dominant kernels (source of roofline.traffic), and the launch list of one batched C5 half sweep.  Numbers printed by a run
# under ncu are never bench values.
mkdir -p gpurun_out
R=${ROUND:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --legs none"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-60000} --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_launches_run.log 2>&1
echo "launch list rc=$? rows=$(wc -l < gpurun_out/${R}_launches.csv)"
python tools/summarize_launches.py gpurun_out/${R}_launches.csv > gpurun_out/${R}_launches_summary.txt; head -n 30 gpurun_out/${R}_launches_summary.txt
for spec in ${NCU_SPECS:-stack_persistent_kernel:60:1 pcg_persistent_kernel:40:1 cholqr_kernel:40:1}; do
  IFS=: read -r kern skip count <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s $skip -c $count \
      -f -o gpurun_out/${R}_prof_$kern $CMD > gpurun_out/${R}_prof_${kern}_run.log 2>&1
  echo "full capture $kern rc=$?"
  ncu -i gpurun_out/${R}_prof_$kern.ncu-rep --page raw --csv > gpurun_out/${R}_prof_${kern}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${R}_prof_$kern.ncu-rep --page source --csv > gpurun_out/${R}_prof_${kern}_source.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${R}_prof_${kern}_raw.csv gpurun_out/${R}_prof_${kern}_source.csv > gpurun_out/${R}_ncu_${kern}.txt 2>/dev/null
done
# batched C5: one launch list (8 systems keep the capture short; the launch COUNT does not depend on the batch size)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"beig|bsvd|gemm|micro_expand|splitk" -c 4000 --csv \
    --log-file gpurun_out/${R}_c5_launches.csv python tools/bench_c5.py 8 8 eigs > gpurun_out/${R}_c5_run.log 2>&1
python - <<PY
import csv, collections, re
rows=[l for l in open("gpurun_out/${R}_c5_launches.csv") if not l.startswith("==")]
agg=collections.OrderedDict()
for rec in csv.DictReader(rows):
    if rec.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(rec["Metric Value"].replace(",","")); u=rec["Metric Unit"]
    us=v/1e3 if u in("nsecond","ns") else v
    k=re.sub(r"\(.*","",rec["Kernel Name"])[:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=us
tot=sum(a[1] for a in agg.values())
with open("gpurun_out/${R}_c5_launches_summary.txt","w") as f:
    f.write("# ncu launch list (gpu__time_duration.sum) of tools/bench_c5.py 8 8 eigs, batched kernels only; serialised, cold cache: compare SHARES\n")
    for k,(n,us) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        f.write(f"{k:<62}{n:>7}{us/1e3:>10.3f} ms{100*us/tot:>7.1f}%{us/n:>10.1f} us\n")
print(open("gpurun_out/${R}_c5_launches_summary.txt").read())
PY
ls -la gpurun_out/*.ncu-rep
