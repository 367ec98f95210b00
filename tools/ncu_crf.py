"""one dense-CRF call on batch 64 @300x300 (for `ncu --set full -k regex:crf_kernel`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_data as synthetic
import mcb200
from mcb200 import postprocessing as G
dev = torch.device("cuda:0")
b = int(os.environ.get("BATCH", "64"))
probs = torch.from_numpy(synthetic.probability_maps(b, 300, seed=7)).to(dev)
img = torch.from_numpy(synthetic.train_batch(b, 300, seed=1)[0]).to(dev)
for _ in range(2):
    out = G.dense_crf_batch(img, probs)
torch.cuda.synchronize()
print("ok", float(out.mean()))
