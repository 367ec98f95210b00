"""DRAM traffic of the tcgen05 GEMM family (conv_gemm_kernel + wgrad_kernel) over ONE train step, from an ncu launch list
taken with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum (tools/r2_final1.sh):
   python tools/gemm_traffic_from_csv.py gpurun_out/r02_launches_ncu.csv > profiles/r02_gemm_traffic.json
The last step = everything from the last stem_im2col launch on."""
import csv, json, sys
from collections import OrderedDict

UNIT = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
BY = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
launch = OrderedDict()
for r in csv.DictReader(lines):
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    m, u = r["Metric Name"], r.get("Metric Unit", "")
    if m.startswith("gpu__time_duration"):
        d["ns"] = v * UNIT.get(u, 1.0)
    elif m.startswith("dram__bytes_read"):
        d["rd"] = v * BY.get(u, 1.0)
    elif m.startswith("dram__bytes_write"):
        d["wr"] = v * BY.get(u, 1.0)
rows = list(launch.values())
start = max(i for i, d in enumerate(rows) if "stem_im2col" in d["name"])
rows = rows[start:]
fam = [d for d in rows if "conv_gemm_kernel" in d["name"] or "wgrad_kernel" in d["name"]]
out = {"source": sys.argv[1], "step_launches": len(rows), "launches": len(fam),
       "dram_read_bytes": int(sum(d.get("rd", 0) for d in fam)), "dram_write_bytes": int(sum(d.get("wr", 0) for d in fam)),
       "time_ms": round(sum(d.get("ns", 0) for d in fam) / 1e6, 3),
       "step_time_ms": round(sum(d.get("ns", 0) for d in rows) / 1e6, 3),
       "step_dram_bytes": int(sum(d.get("rd", 0) + d.get("wr", 0) for d in rows))}
out["dram_bytes"] = out["dram_read_bytes"] + out["dram_write_bytes"]
print(json.dumps(out, indent=1))
