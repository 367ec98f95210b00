"""Diagnostic sweep of the conv GEMM family on the GPU box: prints error summaries instead of asserting."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mcb200
from mcb200 import ops

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
nchw = lambda x: x.permute(0, 3, 1, 2).contiguous()
bf = lambda x: x.to(torch.bfloat16).float()


def report(name, got, ref):
    got = got.float().cpu()
    err = (got - ref).abs()
    rel = err.max() / (ref.abs().max() + 1e-9)
    flag = "OK " if rel < 2e-2 else "BAD"
    print("%s %-40s max_err %.4g ref_max %.4g rel %.3g nan %d" % (flag, name, err.max(), ref.abs().max(), rel,
                                                                    int(torch.isnan(got).sum())), flush=True)
    if flag == "BAD" and got.dim() == 4:
        e = err.amax(dim=(0,))  # C,H,W
        print("   err by channel block of 8:", [round(float(e[c:c + 8].max()), 3) for c in range(0, min(e.shape[0], 64), 8)])
        print("   err by row:", [round(float(e[:, r].max()), 3) for r in range(min(e.shape[1], 16))])
        print("   err by col:", [round(float(e[:, :, r].max()), 3) for r in range(min(e.shape[2], 16))])
        print("   err by image:", [round(float(err[i].max()), 3) for i in range(err.shape[0])])


def run(fn, name):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as ex:
        print("EXC", name, repr(ex)[:300], flush=True)
        traceback.print_exc()


def fwd(n, h, w, cin, cout, k, s):
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(n, cin, h, w, generator=g)); wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** .5)
    ref = F.conv2d(x, wt, None, stride=s, padding=k // 2)
    y = ops.conv_fwd(nhwc(x).to(dev, torch.bfloat16), ops.pack_conv_weight(wt).to(dev, torch.bfloat16), k, s)
    torch.cuda.synchronize()
    report("fwd n%d %dx%d %d->%d k%d s%d" % (n, h, w, cin, cout, k, s), nchw(y), ref)


def dgrad(n, h, w, cin, cout, k, s):
    g = torch.Generator().manual_seed(2)
    wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cout * k * k) ** .5); dy = bf(torch.randn(n, cout, h // s, w // s, generator=g))
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt, dy, stride=s, padding=k // 2)
    dx = ops.conv_dgrad(nhwc(dy).to(dev, torch.bfloat16), ops.pack_conv_weight(wt).to(dev, torch.bfloat16), k, s, (h, w))
    torch.cuda.synchronize()
    report("dgrad n%d %dx%d %d<-%d k%d s%d" % (n, h, w, cin, cout, k, s), nchw(dx), ref)


def wgrad(n, h, w, cin, cout, k, s):
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn(n, cin, h, w, generator=g)); dy = bf(torch.randn(n, cout, h // s, w // s, generator=g))
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, stride=s, padding=k // 2)
    dw = torch.zeros(k * k, cout, cin, device=dev)
    ops.conv_wgrad(nhwc(dy).to(dev, torch.bfloat16), nhwc(x).to(dev, torch.bfloat16), dw, k, s)
    torch.cuda.synchronize()
    report("wgrad n%d %dx%d %d,%d k%d s%d" % (n, h, w, cin, cout, k, s), ops.unpack_conv_weight(dw, k), ref)


def convt(n, h, w, cin, cout):
    g = torch.Generator().manual_seed(4)
    x = bf(torch.randn(n, cin, h, w, generator=g)); wt = bf(torch.randn(cin, cout, 4, 4, generator=g) / (cin * 4) ** .5)
    ref = F.conv_transpose2d(x, wt, None, stride=2, padding=1)
    wp = ops.pack_convt_weight(wt).to(dev, torch.bfloat16)
    xd = nhwc(x).to(dev, torch.bfloat16)
    y = ops.convt_fwd(xd, wp); torch.cuda.synchronize()
    report("convt fwd n%d %dx%d %d->%d" % (n, h, w, cin, cout), nchw(y), ref)
    dy = bf(torch.randn(n, cout, 2 * h, 2 * w, generator=g))
    xr = x.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, None, stride=2, padding=1).backward(dy)
    dyd = nhwc(dy).to(dev, torch.bfloat16)
    dx = ops.convt_dgrad(dyd, wp); torch.cuda.synchronize()
    report("convt dgrad", nchw(dx), xr.grad)
    dw = torch.zeros(16, cout, cin, device=dev)
    ops.convt_wgrad(dyd, xd, dw); torch.cuda.synchronize()
    report("convt wgrad", ops.unpack_convt_weight(dw), wr.grad)


cases = [(2, 16, 16, 64, 64, 1, 1), (2, 16, 16, 64, 128, 3, 1), (2, 16, 16, 128, 256, 3, 1), (2, 16, 16, 64, 128, 3, 2),
         (2, 16, 16, 128, 64, 1, 2), (1, 32, 32, 32, 32, 3, 1), (4, 5, 5, 128, 512, 3, 1), (3, 20, 20, 256, 256, 3, 1)]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for c in cases:
    if which in ("all", "fwd"): run(lambda: fwd(*c), "fwd%s" % (c,))
for c in cases:
    if which in ("all", "dgrad"): run(lambda: dgrad(*c), "dgrad%s" % (c,))
for c in cases:
    if which in ("all", "wgrad"): run(lambda: wgrad(*c), "wgrad%s" % (c,))
for c in [(2, 8, 8, 64, 64), (2, 10, 10, 128, 256), (1, 16, 16, 128, 32)]:
    if which in ("all", "convt"): run(lambda: convt(*c), "convt%s" % (c,))
