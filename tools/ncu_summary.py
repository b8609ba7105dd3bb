"""Key counters of an `ncu --set full` capture, from `ncu -i X.ncu-rep --page raw --csv` (+ optional `--page source --csv`
for the stall-reason totals).   python tools/ncu_summary.py raw.csv [source.csv] > profiles/rNN_ncu_<kernel>.txt"""
import csv, sys, collections

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k, rec in enumerate(rows[2:]):
    print(f"== launch {k}")
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"  {h:<96} {rec[i]} {units[i]}")
if len(sys.argv) > 2:
    srows = list(csv.reader(open(sys.argv[2])))
    shdr = srows[1]
    ix = {h: i for i, h in enumerate(shdr)}
    stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, samples = collections.Counter(), 0
    for r in srows[2:]:
        if len(r) < len(shdr) or not r[ix["# Samples"]].isdigit():
            continue
        samples += int(r[ix["# Samples"]])
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
    print(f"== warp-state samples over the SASS of the kernel: {samples}")
    for s, v in tot.most_common(10):
        print(f"  {s:<28} {v:>8}  {100.0 * v / max(samples, 1):5.1f} %")
