"""Per-kernel summary of an ncu launch list (gpu__time_duration.sum, --csv) of `bench.py --steps 1 --warmup 1`:
the timed step is the span between the 2nd and the 3rd 256 MiB L2-flush fill (every step starts with one).
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import csv, sys, re, collections

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
for rec in csv.DictReader(lines):
    if rec.get("Metric Name") != "gpu__time_duration.sum":
        continue
    unit, val = rec["Metric Unit"], float(rec["Metric Value"].replace(",", ""))
    us = val / 1e3 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1e3)
    rows.append((rec["Kernel Name"], us))
fills = [i for i, (k, _) in enumerate(rows) if "FillFunctor<unsigned char>" in k]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lo, hi = (fills[which], fills[which + 1]) if len(fills) > which + 1 else (0, len(rows))
step = rows[lo:hi]
agg = collections.OrderedDict()
for k, us in step:
    k = re.sub(r"\(.*", "", k)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
print(f"# launches {lo}..{hi - 1} of {len(rows)} = step {which} of the run (between two L2-flush fills); per-launch times under ncu")
print(f"# are cold-cache and serialised: compare SHARES.  total device time in the span: {total / 1e3:.2f} ms, {len(step)} launches")
print(f"{'kernel':<64}{'launches':>9}{'ms':>10}{'share':>8}{'avg us':>10}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:62]:<64}{n:>9}{us / 1e3:>10.3f}{100 * us / total:>7.1f}%{us / n:>10.2f}")
