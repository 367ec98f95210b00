"""One eager pass of every post-processing / instance / TTA kernel on a batch-64 workload (for ncu):
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
       --log-file gpurun_out/postproc_ncu.csv python tools/ncu_postproc.py
summarised by tools/summarize_postproc_ncu.py into profiles/r02_postproc_ncu.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench_data as synthetic
import mcb200
from mcb200 import loaders as lo, postprocessing as G, utils as U

b = int(os.environ.get("BATCH", "64"))
dev = torch.device("cuda:0")
probs = torch.from_numpy(synthetic.probability_maps(b, 320, seed=7)).to(dev)
X = torch.from_numpy(synthetic.train_batch(b, 320, seed=1)[0]).to(dev)
pc = probs[:, :, 10:310, 10:310].contiguous()
img = X[:, :, 10:310, 10:310].contiguous()
torch.cuda.synchronize()
refined = G.dense_crf_batch(img, pc)
pr = G.resize_batch(refined, (300, 300))
masks = G.threshold_batch(pr)
er = G.erode_batch(masks, 2)
lab, cnt = G.label_batch(er, return_counts=True)
dil = G.morph_batch(lab, 2, dilation=True)
sc = G.scores_strided(dil.view(-1, 300, 300), pr.view(-1, 300, 300), cnt)
ws = G.watershed_split(refined[:, 1].contiguous(), hi=0.8, lo=0.5)
arg = G.categorize_batch(refined)
planes = dil.view(-1, 300, 300)
cnts, starts, spans, geo = U.rle_encode_instances(planes, cnt)
specs = lo.tta_specs()
x4 = X[:4].contiguous()
xt = lo.test_time_augmentation_transform_batch(x4, specs * 4, sum(([i] * 16 for i in range(4)), []))
logits = torch.randn(64, 2, 320, 320, device=dev)
agg = lo.aggregate_batch(logits, specs * 4, sum(([i] * 16 for i in range(4)), []), "gmean", from_logits=True)
torch.cuda.synchronize()
print("instances", int(cnt.sum()), "rle counts", len(cnts))
