"""profiles/ncu_traffic.json from an ncu launch list with DRAM byte counters (the file bench.py reads for `traffic`).
usage: python tools/ncu_traffic.py launches.csv syrk_launches_per_step update_path > profiles/ncu_traffic.json"""
import collections
import csv
import json
import sys

path, per_step, upath = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
    v = float(r["Metric Value"].replace(",", ""))
    m, unit = r["Metric Name"], r["Metric Unit"]
    if "bytes" in m:
        agg[name]["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    elif m == "gpu__time_duration.sum":
        agg[name]["launches"] += 1
syrk = ("syrk_i8_persistent_kernel" if "syrk_i8_persistent_kernel" in agg else "syrk_i8_kernel") if upath else "syrk_kernel"
steps = agg[syrk]["launches"] / per_step
chol = sum(agg[k]["bytes"] for k in ("potrf_diag_kernel", "trsm_kernel", "i8_slice_kernel", "i8_rowscale_kernel", syrk)) / steps
out = {
    "source": "%s: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on "
              "`python bench.py --steps 2 --warmup 3 --no-cpu-baseline`, first 2500 launches = %.2f steps, 1 x B200" % (path, steps),
    "workload": "C3", "walkers": 128, "update_path": upath,
    "cholesky_dram_bytes_per_step": chol,
    "syrk_dram_bytes_per_step": agg[syrk]["bytes"] / steps,
    "trsm_dram_bytes_per_step": agg["trsm_kernel"]["bytes"] / steps,
    "slice_dram_bytes_per_step": agg["i8_slice_kernel"]["bytes"] / steps,
    "assembly_dram_bytes_per_launch": agg["assemble_fast_kernel"]["bytes"] / max(agg["assemble_fast_kernel"]["launches"], 1),
    "note": "dram__bytes_read.sum + dram__bytes_write.sum; ncu replays each kernel from a flushed cache and serialises the "
            "sub-batch streams, so operand panels that partly hit in L2 in the real run are counted in full",
}
print(json.dumps(out, indent=1))
