"""where does the end-to-end arm lose time against the device-resident arm? (gpurun diagnostic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mcb200
from mcb200.models import PyTorchUNetWeighted
from oracle import synthetic

dev = torch.device("cuda:0")
m = PyTorchUNetWeighted(**bench.unet_config("ResNet101")); m._to_device()
x, t = synthetic.train_batch(32, 320, seed=1)
Xh, Th = torch.from_numpy(x).pin_memory(), torch.from_numpy(t).pin_memory()
Xd, Td = Xh.to(dev), Th.to(dev)
for _ in range(3):
    m._fit_loop([Xd, Td]); m._fit_loop([Xh, Th])
torch.cuda.synchronize()


def timed(fn, n=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def host_only(fn, n=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return dt


sx, st = torch.empty_like(Xd), torch.empty_like(Td)
print("H2D X+T alone          %.3f ms" % timed(lambda: (sx.copy_(Xh, non_blocking=True), st.copy_(Th, non_blocking=True))))
print("device inputs          %.3f ms/step (host enqueue %.3f)" % (timed(lambda: m._fit_loop([Xd, Td])), host_only(lambda: m._fit_loop([Xd, Td]))))
print("host inputs, no read   %.3f ms/step (host enqueue %.3f)" % (timed(lambda: m._fit_loop([Xh, Th])), host_only(lambda: m._fit_loop([Xh, Th]))))
prev = {"l": None}
def e2e():
    cur = m._fit_loop([Xh, Th])["sum"]
    if prev["l"] is not None:
        float(prev["l"].cpu())
    prev["l"] = cur
print("host inputs + loss read %.3f ms/step" % timed(e2e))
def e2e_sync():
    float(m._fit_loop([Xh, Th])["sum"].cpu())
print("host inputs + sync read %.3f ms/step" % timed(e2e_sync))
