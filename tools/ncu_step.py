"""Runs a few fused train steps of the bench workload (for ncu): 1 eager step, graph capture, N replays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mcb200
from mcb200.models import PyTorchUNetWeighted
from oracle import synthetic

enc = int(os.environ.get("ENC", "101")); b = int(os.environ.get("BATCH", "32")); s = int(os.environ.get("SIZE", "320"))
replays = int(os.environ.get("REPLAYS", "2"))
torch.manual_seed(1234)
m = PyTorchUNetWeighted(**bench.unet_config("ResNet%d" % enc)); m._to_device()
x, t = synthetic.train_batch(b, s, seed=1234)
X, T = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
for i in range(1 + replays):
    loss = m._fit_loop([X, T])["sum"]
    torch.cuda.synchronize()
    print("step", i, float(loss), flush=True)
print("launches per step (plan):", m._fused.count_launches())
