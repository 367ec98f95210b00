"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> markdown table per
kernel: launches, time, DRAM bytes, achieved DRAM GB/s against the measured HBM peak.
usage: summarize_postproc_ncu.py gpurun_out/postproc_ncu.csv [batch] > profiles/r02_postproc_ncu.md"""
import csv, json, os, re, sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6567.7
pj = os.path.join(root, "MEASURED_PEAKS.json")
if os.path.exists(pj):
    peak = json.load(open(pj)).get("hbm_gbs", peak)
UNIT = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
BY = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
launch = OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    d = launch.setdefault(key, {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    m, u = r["Metric Name"], r.get("Metric Unit", "")
    if m.startswith("gpu__time_duration"):
        d["ns"] = v * UNIT.get(u, 1.0)
    elif m.startswith("dram__bytes_read"):
        d["rd"] = v * BY.get(u, 1.0)
    elif m.startswith("dram__bytes_write"):
        d["wr"] = v * BY.get(u, 1.0)


def fam(name):
    m = re.match(r"(?:void )?(?:mcb::)?([A-Za-z0-9_]+)", name)
    return m.group(1) if m else name


acc = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in launch.values():
    if "at::" in d["name"] or "elementwise_kernel" in d["name"]:
        k = "(torch helper) " + fam(d["name"])
    else:
        k = fam(d["name"])
    a = acc[k]
    a[0] += 1
    a[1] += d.get("ns", 0.0)
    a[2] += d.get("rd", 0.0)
    a[3] += d.get("wr", 0.0)
print("# post-processing / instance / TTA kernels, batch %d @300x300: ncu per-kernel device time and DRAM traffic\n" % batch)
print("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over "
      "`tools/ncu_postproc.py` (cold-cache, serialised replay: durations are upper bounds of the in-graph cost). "
      "HBM peak = %.0f GB/s (MEASURED_PEAKS.json).\n" % peak)
print("| kernel | launches | total us | DRAM read MB | DRAM write MB | DRAM GB/s | of HBM peak |\n|---|---|---|---|---|---|---|")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / a[1] if a[1] else 0.0
    print("| `%s` | %d | %.1f | %.2f | %.2f | %.0f | %.1f%% |" % (k, a[0], a[1] / 1e3, a[2] / 1e6, a[3] / 1e6, gbs, 100 * gbs / peak))
