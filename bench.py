#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path (contract in the task statement, section 4).

  python bench.py --gpus N --steps K --warmup W          our arm: ResNet101-UNet train step, batch 32/GPU, 300x300
                                                          tiles replicate-padded to the 320x320 net input
  python bench.py --impl reference ...                    the reference's CPU path (oracle port) on the host cores

One JSON line on stdout (rank 0).  `value` = tiles/s with inputs resident in HBM; `e2e` = the same metric through the
reference-facing API (PyTorchUNetWeighted._fit_loop on pinned HOST batches, loss read back every step)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TILE = {(101, 320): 205.71e9, (101, 256): 131.7e9, (34, 256): 79.54e9, (152, 512): 643.0e9}


def unet_config(encoder, lr=5e-4):
    """the `config.unet` dict of /root/reference/src/pipeline_config.py:61-120 (values from neptune.yaml)"""
    return {
        'architecture_config': {
            'model_params': {'n_filters': 16, 'conv_kernel': 3, 'pool_kernel': 3, 'pool_stride': 2, 'repeat_blocks': 4,
                             'batch_norm': 1, 'dropout': 0.1, 'in_channels': 3, 'out_channels': 2, 'nr_outputs': 1,
                             'encoder': encoder},
            'optimizer_params': {'lr': lr},
            'regularizer_params': {'regularize': True, 'weight_decay_conv2d': 1e-4},
            'weights_init': {'function': 'he'},
            'loss_weights': {'bce_mask': 1.0, 'dice_mask': 0.2},
            'weighted_cross_entropy': {'w0': 50, 'sigma': 10, 'imsize': (256, 256)},
            'dice': {'smooth': 1, 'dice_activation': 'softmax'},
        },
        'training_config': {'epochs': 1},
        'callbacks_config': {},
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7:
                try:
                    rows.append((float(f[0]), float(f[1]), float(f[2]), f[3], f[4], f[5], f[6]))
                except ValueError:
                    pass
        os.unlink(self.path)
        if not rows:
            return out
        busy = [r for r in rows if r[2] > 300] or rows
        sm = sorted(r[0] for r in busy)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = rows[0][1]
        out["power_w_max"] = max(r[2] for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return out


def usable_cores():
    """host cores this process may actually use: affinity mask, cgroup CPU quota, capped at 32 (torch's intra-op
    parallelism stops scaling — and on over-subscribed boxes collapses — beyond that for these layer sizes)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        pass
    return max(1, min(n, 32))


def cpu_reference_arm(args, sample_batch=None, steps=None, warmup=None):
    """the reference's own CPU implementation of the train step (oracle port of Model._fit_loop, fp32, all host
    threads) on a bounded sample of the workload"""
    import torch
    from oracle import unet_oracle as O, synthetic
    cores = usable_cores()
    torch.set_num_threads(cores)
    b = sample_batch or max(1, min(args.batch, 4))
    steps = steps or max(1, min(args.steps, 3))
    warmup = 1 if warmup is None else warmup
    sd = O.make_reference_like_state_dict(args.encoder, seed=1234)
    x, t = synthetic.train_batch(b, args.size, seed=1234)
    X, T = torch.from_numpy(x), torch.from_numpy(t)
    opt = O.AdamOracle(lr=5e-4, weight_decay=1e-4)
    tw = time.time()
    for _ in range(warmup):
        O.train_step(sd, args.encoder, X, T, opt, imsize=(256, 256))
    tw = (time.time() - tw) / max(warmup, 1)
    if warmup and tw * steps > 60:  # keep the whole arm bounded
        steps = max(1, int(60 / tw))
    t0 = time.time()
    for _ in range(steps):
        O.train_step(sd, args.encoder, X, T, opt, imsize=(256, 256))
    dt = (time.time() - t0) / steps
    return {"value": b / dt, "unit": "tiles/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle port of Model._fit_loop (fp32 torch CPU), UNetResNet-%d, batch %d @%dx%d, %d warm-up + %d timed steps"
                      % (args.encoder, b, args.size, args.size, warmup, steps), "ms_per_step": dt * 1e3}


def breakdown(step, n_iter=2):
    """instrumented eager pass: CUDA events around every launch of the plan, grouped by kernel family"""
    import torch
    plan = step.plan
    ops = [("fwd", o) for o in plan.fwd_ops] + [("bwd", o) for l in plan.bwd_layers for o in l]
    acc = {}
    for it in range(n_iter):
        plan._stats_arena.zero_()
        evs = []
        for phase, o in ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o()
            e1.record()
            evs.append((o, e0, e1))
        torch.cuda.synchronize()
        if it == n_iter - 1:
            for o, e0, e1 in evs:
                a = acc.setdefault(o.kind, [0, 0.0, 0.0, 0.0])
                a[0] += 1
                a[1] += e0.elapsed_time(e1)
                a[2] += o.flops
                a[3] += o.bytes
    total = sum(a[1] for a in acc.values())
    per_op = os.environ.get("MCB_BENCH_PER_OP")
    if per_op:
        rows = sorted(((e0.elapsed_time(e1), o) for o, e0, e1 in evs), key=lambda r: -r[0])
        with open(per_op, "w") as f:
            for ms, o in rows:
                f.write("%8.4f ms  %-14s %-40s %7.1f TFLOP/s %8.1f GB/s\n" % (
                    ms, o.kind, o.desc, o.flops / (ms * 1e-3) / 1e12 if ms > 0 else 0,
                    o.bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0))
    out = {}
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        out[k] = {"launches": a[0], "ms": round(a[1], 3), "share": round(a[1] / total, 4),
                  "tflops": round(a[2] / (a[1] * 1e-3) / 1e12, 1) if a[2] else None,
                  "algo_gbs": round(a[3] / (a[1] * 1e-3) / 1e9, 1) if a[3] else None}
    return out, total


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else this process (or NCCL's banner, printed from C) writes
    to fd 1 has been rerouted to stderr by main()"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--encoder", type=int, default=101)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = "UNetResNet-%d train step (fwd + weighted-CE/Dice loss + bwd + Adam), batch %d/GPU, 300x300 tiles " \
               "replicate-padded to the %dx%d net input (reference loader_mode crop_and_pad)" % (
                   args.encoder, args.batch, args.size, args.size)

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(args)
        line = {"impl": "reference", "metric": "300x300 tiles/sec fwd+bwd ResNet101-UNet", "value": cb["value"],
                "unit": "tiles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "global_batch": args.batch * args.gpus, "parallelism": "cpu"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import mcb200
    from mcb200.models import PyTorchUNetWeighted
    from oracle import synthetic

    dev = torch.device("cuda", local_rank)
    t_start = time.time()

    def note(msg):
        if os.environ.get("MCB_BENCH_VERBOSE"):
            sys.stderr.write("[bench %7.2fs] %s\n" % (time.time() - t_start, msg))
            sys.stderr.flush()

    torch.manual_seed(1234)
    model = PyTorchUNetWeighted(**unet_config("ResNet%d" % args.encoder))
    model._to_device()
    x, t = synthetic.train_batch(args.batch, args.size, seed=1234 + rank)
    Xh, Th = torch.from_numpy(x).pin_memory(), torch.from_numpy(t).pin_memory()
    Xd, Td = Xh.to(dev), Th.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # ---- device-resident arm
    last = {}

    def step_dev():
        last["loss"] = model._fit_loop([Xd, Td])["sum"]

    note("model + data ready")
    for i in range(args.warmup):
        step_dev()
        torch.cuda.synchronize()
        note("warm-up step %d done" % i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(step_dev, args.steps)
    clocks = sampler.stop() if sampler else None
    loss_dev = float(last["loss"])
    note("device arm timed: %.3f ms/step" % ms_dev)

    # ---- end-to-end arm: pinned host batches in, loss read back EVERY step.  One step of pipelining: the D2H copy of
    # step i's loss is enqueued (into pinned memory, followed by an event) right behind step i, and the host consumes it
    # after it has enqueued step i+1 -- so the H2D copy of the next batch and the host-side launch work overlap compute.
    # (A plain loss.cpu() of the previous step would be stream-ordered behind the step just enqueued and stall the host
    # for a whole step: tools/dbg_e2e.py.)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    prev = {"i": 0, "pending": False}

    def step_e2e():
        i = prev["i"]
        cur = model._fit_loop([Xh, Th])["sum"]
        loss_host[i & 1].copy_(cur, non_blocking=True)
        loss_ev[i & 1].record()
        if prev["pending"]:
            loss_ev[(i - 1) & 1].synchronize()
            prev["val"] = float(loss_host[(i - 1) & 1])
        prev["pending"] = True
        prev["i"] = i + 1

    for _ in range(args.warmup):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    note("e2e arm timed: %.3f ms/step" % ms_e2e)

    # data-parallel sanity: every rank trains on different tiles, so the replicas stay identical only if the gradient
    # all-reduce really happened -- a parameter checksum must agree across ranks (a step without the exchange would be
    # faster and meaningless)
    in_sync = None
    if world > 1:
        chk = model._net()._p32.double().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        vals = [float(c) for c in allc]
        in_sync = (max(vals) - min(vals)) <= 1e-9 * max(1.0, abs(vals[0]))
        if not in_sync:
            raise RuntimeError("replicas diverged (parameter checksums %r): gradient all-reduce missing?" % (vals,))

    tiles = args.batch * world
    value = tiles / (ms_dev * 1e-3)
    e2e = tiles / (ms_e2e * 1e-3)
    fused = model._fused
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, peak_src = peaks()
    fpt = FLOP_PER_TILE.get((args.encoder, args.size))
    if fpt is None:
        fpt = (sum(o.flops for o in fused.plan.fwd_ops) + sum(o.flops for l in fused.plan.bwd_layers for o in l)) / args.batch
    achieved = value / world * fpt / 1e12
    line = {
        "metric": "300x300 tiles/sec fwd+bwd ResNet101-UNet", "value": round(value, 2), "unit": "tiles/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": tiles, "parallelism": "dp%d" % world,
                   "bn": "per-replica batch statistics (reference DataParallel semantics)" if world > 1 else "batch statistics",
                   "replicas_in_sync": in_sync,
                   "l2": "per-step working set (activations + gradients, GBs) far exceeds the 126 MB L2; no flush needed",
                   "loss_last_step": loss_dev},
        "e2e": {"value": round(e2e, 2), "unit": "tiles/s", "ms_per_step": round(ms_e2e, 4),
                "h2d_bytes_per_step": int(Xh.numel() * 4 + Th.numel() * 4), "d2h_bytes_per_step": 4,
                "api": "mcb200.models.PyTorchUNetWeighted._fit_loop([X_host_pinned, target_host_pinned]) + async D2H of the loss into pinned memory, consumed one step later"},
        "gpu_launches": fused.count_launches() * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(achieved / peak_tf, 4), "traffic": None, "peak_source": peak_src,
                     "what": "conv/convT fwd+dgrad+wgrad tcgen05 GEMM family: algorithmic conv FLOPs of the step "
                             "(%.2f GFLOP/tile) / whole-step time per GPU" % (fpt / 1e9)},
    }
    # DRAM traffic of the GEMM family over one step, from the committed ncu capture of this same workload (a number
    # taken under the profiler is evidence, not a bench value): compare with the family's algorithmic bytes
    tj = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01c_gemm_traffic.json")
    if (args.encoder, args.batch, args.size) == (101, 32, 320) and os.path.exists(tj):
        with open(tj) as f:
            tr = json.load(f)
        line["roofline"]["traffic"] = int(tr["dram_bytes"])
        line["roofline"]["traffic_unit"] = "DRAM bytes per step per GPU, summed over the family's %d launches " \
                                           "(ncu dram__bytes_read.sum + dram__bytes_write.sum, cold-cache replay)" % tr["launches"]
        line["roofline"]["traffic_source"] = "profiles/r01c_gemm_traffic.json"
    if not args.no_breakdown:
        bd, total = breakdown(fused)
        note("breakdown done")
        line["breakdown"] = bd
        gemm = [v for k, v in bd.items() if k.startswith("conv")]
        gemm_ms = sum(v["ms"] for v in gemm)
        gemm_fl = sum((v["tflops"] or 0) * v["ms"] for v in gemm)
        line["roofline"]["gemm_family_ms"] = round(gemm_ms, 3)
        line["roofline"]["gemm_family_tflops"] = round(gemm_fl / gemm_ms, 1) if gemm_ms else None
        line["roofline"]["gemm_family_share_of_step"] = round(gemm_ms / total, 4)
        line["roofline"]["gemm_family_algorithmic_bytes"] = int(sum((v["algo_gbs"] or 0) * v["ms"] * 1e6 for v in gemm))
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_arm(args, sample_batch=2, steps=2, warmup=1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        note("cpu baseline done")
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
