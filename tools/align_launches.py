"""Aligns an ncu launch list (gpu__time_duration per launch) of ONE train step with the plan's op list, giving the
device time, achieved TFLOP/s and algorithmic GB/s of every layer.  usage:
   python tools/align_launches.py launches.csv [encoder batch size] > profiles/xxx_per_layer.md"""
import csv, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcb200
from mcb200.unet_models import UNetResNet

path = sys.argv[1]
enc, batch, size = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (101, 32, 320)
KERNEL = {"conv_fwd": "conv_gemm_kernel", "convt_fwd": "conv_gemm_kernel", "conv_dgrad": "conv_gemm_kernel",
          "convt_dgrad": "conv_gemm_kernel", "conv_wgrad": "wgrad_kernel", "convt_wgrad": "wgrad_kernel",
          "bn_apply": "bn_train_apply_kernel", "bn_bwd_reduce": "channel_reduce_kernel<1>",
          "bn_bwd_apply": "bn_bwd_apply_kernel", "channel_sum": "channel_reduce_kernel<0>", "maxpool": "maxpool2_",
          "stem_im2col": "stem_im2col_kernel", "final_conv": "final_conv_"}
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name", "").startswith("gpu__time_duration"):
        rows.append((r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", ""))))
# last step = from the last stem_im2col launch onwards
start = max(i for i, r in enumerate(rows) if "stem_im2col" in r[0])
rows = rows[start:]
net = UNetResNet(enc, 2, 32, 0.0, False, True)
plan = net.plan(batch, size, size, True) if False else None
# build the plan on the meta-free CPU path with batch 1 and scale costs (allocating batch 32 on CPU is wasteful)
plan = net.plan(1, size, size, True)
ops = list(plan.fwd_ops) + [o for l in plan.bwd_layers for o in l]
out = []
# one queue per kernel: the weight-gradient GEMMs run on a side stream (launched later than their place in the plan), but
# the launches of ONE kernel keep their plan order
queues = {want: [r for r in rows if want in r[0]] for want in set(KERNEL.values())}
ptr = {want: 0 for want in queues}
for o in ops:
    want = KERNEL.get(o.kind)
    if want is None:
        continue
    if ptr[want] >= len(queues[want]):
        continue
    name, grid, ns = queues[want][ptr[want]]
    ptr[want] += 1
    m = re.search(r"(conv_gemm_kernel|wgrad_kernel)(<[^>]*>)", name)
    out.append((ns, o.kind, o.desc.replace("@1x", "@%dx" % batch), (m.group(2) if m else ""), grid, o.flops * batch, o.bytes * batch))
tot = sum(r[0] for r in out)
print("# per-layer device time of one train step (UNetResNet-%d, batch %d, %dx%d) — ncu gpu__time_duration, cold-cache, serialised\n" % (enc, batch, size, size))
print("aligned %d launches, %.3f ms\n" % (len(out), tot / 1e6))
print("| ms | kind | layer | tile cfg | grid | TFLOP/s | algorithmic GB/s |\n|---|---|---|---|---|---|---|")
for ns, kind, desc, cfg, grid, fl, by in sorted(out, key=lambda r: -r[0])[: int(os.environ.get("TOP", "70"))]:
    print("| %.4f | %s | %s | %s | %s | %s | %.0f |" % (ns / 1e6, kind, desc, cfg, grid, ("%.0f" % (fl / ns / 1e3)) if fl else "-", by / ns))
# aggregate by (kind, desc)
agg = {}
for ns, kind, desc, cfg, grid, fl, by in out:
    a = agg.setdefault((kind, desc, cfg), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += ns; a[2] += fl; a[3] += by
print("\n## aggregated by layer shape\n\n| total ms | n | kind | layer | cfg | TFLOP/s | GB/s |\n|---|---|---|---|---|---|---|")
for (kind, desc, cfg), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("| %.3f | %d | %s | %s | %s | %s | %.0f |" % (a[1] / 1e6, a[0], kind, desc, cfg, ("%.0f" % (a[2] / a[1] / 1e3)) if a[2] else "-", a[3] / a[1]))
