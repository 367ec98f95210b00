"""Diagnostic: GPU post-processing vs the CPU oracle, with timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcb200
from mcb200 import postprocessing as G
from oracle import post_oracle as P, synthetic

dev = torch.device("cuda:0")
N, S, T = 8, 256, 300
probs = synthetic.probability_maps(N, S, seed=3)
pd = torch.from_numpy(probs).to(dev)
r = G.resize_batch(pd, (T, T)); torch.cuda.synchronize()
ref_r = np.stack([P.resize_image(p, (T, T)) for p in probs])
print("resize bit-exact:", np.array_equal(r.cpu().numpy(), ref_r), "maxdiff", np.abs(r.cpu().numpy() - ref_r).max())
m = G.threshold_batch(r); ref_m = np.stack([P.categorize_multilayer_image(x) for x in ref_r])
print("threshold exact:", np.array_equal(m.cpu().numpy().astype(bool), ref_m))
lab, cnt = G.label_batch(m, return_counts=True); ref_l = np.stack([P.label_multilayer_image(x) for x in ref_m])
print("label exact:", np.array_equal(lab.cpu().numpy(), ref_l), "K", cnt.cpu().numpy().tolist()[:6], ref_l.reshape(N * 2, -1).max(1).tolist()[:6])
for k in (1, 2, 3, 4, 5):
    d = G.morph_batch(lab, k, True); ref_d = np.stack([P.dilate_image(x, k) for x in ref_l])
    e = G.erode_batch(m, k); ref_e = np.stack([P.erode_image(x.astype(np.uint8) != 0, k) for x in ref_m])
    print("k=%d dilate exact: %s  erode+add_dropped exact: %s" % (k, np.array_equal(d.cpu().numpy(), ref_d), np.array_equal(e.cpu().numpy(), ref_e)))
d2 = G.morph_batch(lab, 2, True)
sc, offs, cnts = G.scores_batch(d2.view(-1, T, T), r.view(-1, T, T), cnt)
ref_sc = []
for i in range(N):
    _, s = P.build_score(P.dilate_image(ref_l[i], 2), ref_r[i])
    ref_sc += [float(v) if v is not np.ma.masked else np.nan for layer in s for v in layer]
got = sc.cpu().numpy(); ref_sc = np.array(ref_sc)
ok = np.allclose(got, ref_sc, rtol=1e-9, atol=0, equal_nan=True)
print("scores match:", ok, len(got), len(ref_sc), np.nanmax(np.abs(got - ref_sc) / np.abs(ref_sc)))
# full transformer timing, batch 64
probs64 = torch.from_numpy(synthetic.probability_maps(64, S, seed=5)).to(dev)
pp = G.MaskPostprocessor((T, T), "resize", 0, 2)
for _ in range(3): pp.run_device(probs64)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(10): out = pp.run_device(probs64)
torch.cuda.synchronize(); print("postproc batch 64: %.3f ms/batch device-resident" % ((time.time() - t0) / 10 * 1e3))
t0 = time.time()
for i in range(8):
    x = P.resize_image(probs[i], (T, T)); c = P.categorize_multilayer_image(x); l = P.label_multilayer_image(c); dd = P.dilate_image(l, 2); P.build_score(dd, x)
print("cpu oracle chain: %.1f ms/image" % ((time.time() - t0) / 8 * 1e3))
