"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
usage: python tools/ncu_summary.py launches.csv [skip_launches | last-step]"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v))
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
if skip == "last-step":
    # the step starts with sample_timesteps' randint (the only uint32 distribution kernel of a step)
    marks = [i for i, (k, _) in enumerate(rows) if "distribution_elementwise_grid_stride_kernel<unsigned int" in k]
    rows = rows[marks[-1]:] if marks else rows
else:
    rows = rows[int(skip):]
tot = sum(v for _, v in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1; agg[k][1] += v
print(f"{len(rows)} launches, {tot/1e3:.2f} ms total kernel time")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v/1e3:9.3f} ms {100*v/tot:5.1f}%  x{n:<5d} {k[:100]}")
