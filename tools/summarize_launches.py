"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-family count / time / share for the
LAST step (last `--per-step` launches).  usage: summarize_launches.py launches.csv [per_step] > profiles/xxx.md"""
import csv, re, sys
from collections import defaultdict

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name", "").startswith("gpu__time_duration"):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6, "second": 1e9, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""), ns))
if per_step:
    rows = rows[-per_step:]


def family(name):
    m = re.match(r"(?:void )?(?:mcb::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    base = m.group(1) if m else name
    targs = m.group(2) or "" if m else ""
    return base + targs


acc = defaultdict(lambda: [0, 0.0])
for name, grid, block, ns in rows:
    a = acc[family(name)]
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in acc.values())
print("# launch list summary: %s (%d launches, %.3f ms serialised GPU time)\n" % (path, len(rows), total / 1e6))
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k, a[0], a[1] / 1e6, 100 * a[1] / total, a[1] / a[0] / 1e3))
print("\n## 40 longest launches\n\n| kernel | grid | ms |\n|---|---|---|")
for name, grid, block, ns in sorted(rows, key=lambda r: -r[3])[:40]:
    print("| `%s` | %s | %.4f |" % (family(name), grid, ns / 1e6))
