"""Micro-benchmark sweep of the conv GEMM family on the bench shapes under different library tuning env vars.
Each op is captured 8x into a CUDA graph and replayed (no host overhead in the numbers)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcb200
from mcb200 import ops

dev = torch.device("cuda:0")
BF = torch.bfloat16
N = 32


def timed(fn, reps=8, replays=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e3  # us


def mk(*shape, dtype=BF):
    return (torch.randn(*shape, device=dev) * 0.1).to(dtype)


def fwd(cin, cout, k, hw, stats=True):
    x, w = mk(N, hw, hw, cin), mk(k * k, cout, cin)
    y = torch.empty(N, hw, hw, cout, dtype=BF, device=dev)
    st = torch.zeros(2 * cout, device=dev) if stats else None
    fl = 2.0 * N * hw * hw * cin * cout * k * k
    return (lambda: ops.conv_fwd(x, w, k, 1, stats=st, out=y)), fl


def dgrad(cin, cout, k, hw):
    dy, w = mk(N, hw, hw, cout), mk(k * k, cout, cin)
    dx = torch.empty(N, hw, hw, cin, dtype=BF, device=dev)
    fl = 2.0 * N * hw * hw * cin * cout * k * k
    return (lambda: ops.conv_dgrad(dy, w, k, 1, (hw, hw), out=dx)), fl


def dgrad_fused(cin, cout, k, hw):
    """the backward's real epilogue: ReLU mask of the producer + its BatchNorm-backward reductions"""
    dy, w = mk(N, hw, hw, cout), mk(k * k, cout, cin)
    dx = torch.empty(N, hw, hw, cin, dtype=BF, device=dev)
    z = mk(N, hw, hw, cin)
    mean, invstd = torch.zeros(cin, device=dev), torch.ones(cin, device=dev)
    gamma, beta = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
    dbeta, dgamma = torch.zeros(cin, device=dev), torch.zeros(cin, device=dev)
    fl = 2.0 * N * hw * hw * cin * cout * k * k
    return (lambda: ops.conv_dgrad(dy, w, k, 1, (hw, hw), out=dx,
                                   bn_reduce=(z, mean, invstd, gamma, beta, dbeta, dgamma))), fl


def dgrad_mask(cin, cout, k, hw):
    dy, w = mk(N, hw, hw, cout), mk(k * k, cout, cin)
    dx = torch.empty(N, hw, hw, cin, dtype=BF, device=dev)
    y = mk(N, hw, hw, cin)
    fl = 2.0 * N * hw * hw * cin * cout * k * k
    return (lambda: ops.conv_dgrad(dy, w, k, 1, (hw, hw), out=dx, relu_mask=y)), fl


def wgrad(cin, cout, k, hw):
    dy, x = mk(N, hw, hw, cout), mk(N, hw, hw, cin)
    dw = torch.zeros(k * k, cout, cin, device=dev)
    fl = 2.0 * N * hw * hw * cin * cout * k * k
    return (lambda: ops.conv_wgrad(dy, x, dw, k, 1)), fl


SHAPES = [(256, 1024, 1, 20), (1024, 256, 1, 20), (256, 256, 3, 20), (128, 128, 3, 160), (32, 32, 3, 320),
          (64, 256, 1, 80), (512, 512, 3, 10), (128, 128, 3, 40), (64, 64, 3, 80), (256, 64, 1, 80)]
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(kind, maker, envs):
    for shp in SHAPES:
        fn, fl = maker(*shp)
        row = []
        for env in envs:
            for k, v in env.items():
                os.environ[k] = str(v)
            try:
                us = timed(fn)
                row.append("%7.1fus %5.0fTF" % (us, fl / us / 1e6))
            except Exception as ex:
                row.append("ERR %s" % str(ex)[:30])
            for k in env:
                os.environ.pop(k, None)
        print("%-6s %4d->%4d k%d @%3d | " % ((kind,) + shp) + " | ".join(row), flush=True)


if which in ("all", "fwd"):
    envs = [{}, {"MCB_HALO": 1}]
    print("fwd   envs:", envs)
    run("fwd", fwd, envs)
    run("dgrad", dgrad, envs)
    run("dgradF", dgrad_fused, envs)
    run("dgradM", dgrad_mask, envs[:1])
if which in ("all", "wgrad"):
    envs = [{}, {"MCB_WGRAD_WAVES_X10": 5}, {"MCB_WGRAD_KB_TARGET": 32}, {"MCB_WGRAD_KB_TARGET": 64},
            {"MCB_WGRAD_KB_TARGET": 128}, {"MCB_WGRAD_KB_TARGET": 64, "MCB_WGRAD_MIN_WAVE_X10": 3},
            {"MCB_WGRAD_KB_TARGET": 128, "MCB_WGRAD_MIN_WAVE_X10": 3}, {"MCB_WGRAD_KB_TARGET": 128, "MCB_WGRAD_MIN_WAVE_X10": 2}]
    print("wgrad envs:", envs)
    run("wgrad", wgrad, envs)
