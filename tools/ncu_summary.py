"""Digest an `ncu --csv --metrics ...` launch list into a per-kernel markdown table.
usage: python tools/ncu_summary.py launches.csv [launches_per_step] > summary.md"""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0]
    m = r["Metric Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
    elif "bytes" in m:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d = agg.setdefault(name, collections.defaultdict(float))
    d[m] += v
    if m == "gpu__time_duration.sum":
        d["launches"] += 1
tot = sum(d["gpu__time_duration.sum"] for d in agg.values())
print("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |")
print("|---|---|---|---|---|---|")
for name, d in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    print("| `%s` | %d | %.3f | %.1f %% | %s | %s |" % (
        name, d["launches"], d["gpu__time_duration.sum"], 100 * d["gpu__time_duration.sum"] / tot,
        ("%.3f" % (d["dram__bytes_read.sum"] / 1e9)) if "dram__bytes_read.sum" in d else "-",
        ("%.3f" % (d["dram__bytes_write.sum"] / 1e9)) if "dram__bytes_write.sum" in d else "-"))
print("\ntotal kernel time %.3f ms over %d launches" % (tot, sum(d["launches"] for d in agg.values())))
