"""BASELINE.json configs[3]: inference-only ResNet101-UNet + full post-processing, batch 64, 300x300 tiles, 1 GPU.
Reports the forward time, the post-processing time (threshold -> [CRF] -> erode -> label -> dilate -> score
[-> watershed]) and post-processing as a share of the step.  Everything stays on the device; the only D2H transfers
are the per-plane instance counts (256 B) inside the score step and the final labels/scores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import mcb200
from mcb200 import ops, postprocessing as G
from mcb200.models import PyTorchUNet
from oracle import synthetic

enc = int(os.environ.get("ENC", "101")); b = int(os.environ.get("BATCH", "64")); s = int(os.environ.get("SIZE", "320"))
dev = torch.device("cuda:0")
torch.manual_seed(1234)
m = PyTorchUNet(**bench.unet_config("ResNet%d" % enc)); m._to_device()
net = m.model; net.eval()
x, _ = synthetic.train_batch(b, s, seed=1234)
X = torch.from_numpy(x).to(dev)
# realistic probability maps for the post-processing (a random-init net predicts noise): synthetic buildings
probs_syn = torch.from_numpy(synthetic.probability_maps(b, s, seed=7)).to(dev)
mode = "crop" if s == 320 else "resize"
pp = G.MaskPostprocessor((300, 300), mode, erode_selem_size=2, dilate_selem_size=2)


def ev(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_fwd = ev(lambda: ops.softmax2(net(X)))
t_pp_eager = ev(lambda: pp.run_device(probs_syn))
t_pp = ev(lambda: pp.run_device_graphed(probs_syn))
pp_default = G.MaskPostprocessor((300, 300), mode, erode_selem_size=0, dilate_selem_size=0)  # neptune.yaml:69-70 defaults
t_pp_default_eager = ev(lambda: pp_default.run_device(probs_syn))
t_pp_default = ev(lambda: pp_default.run_device_graphed(probs_syn))
crop = probs_syn[:, :, 10:310, 10:310].contiguous() if s == 320 else G.resize_batch(probs_syn, (300, 300)).float()
img = X[:, :, 10:310, 10:310].contiguous() if s == 320 else torch.randn(b, 3, 300, 300, device=dev)
t_crf = ev(lambda: G.dense_crf_batch(img, crop), n=3, warm=1)
p1 = crop[:, 1].contiguous()
t_ws = ev(lambda: G.watershed_split(p1, hi=0.8, lo=0.5), n=3, warm=1)
stages = {}
pr = crop
stages["threshold"] = ev(lambda: G.threshold_batch(pr))
masks = G.threshold_batch(pr)
stages["erode+add_dropped(2)"] = ev(lambda: G.erode_batch(masks, 2))
stages["label"] = ev(lambda: G.label_batch(masks))
lab, cnt = G.label_batch(masks, return_counts=True)
stages["dilate(2)"] = ev(lambda: G.morph_batch(lab, 2, True))
stages["score"] = ev(lambda: G.scores_strided(lab.view(-1, 300, 300), pr.reshape(-1, 300, 300), cnt))
if mode == "resize":
    stages["resize"] = ev(lambda: G.resize_batch(probs_syn, (300, 300)))


def cpu_postproc_ms_per_image(probs, n_img=8):
    """the reference's own serial per-image chain (src/pipelines.py:248-304 through src/utils.py:352-355) restated by
    the oracle, on the host CPU: crop/resize -> categorize -> [erode] -> label -> dilate -> build_score"""
    from oracle import post_oracle as P
    imgs = probs[:n_img].cpu().numpy()
    t0 = time.perf_counter()
    for p_ in imgs:
        r = P.crop_image_center_per_class(p_, 300, 300) if mode == "crop" else P.resize_image(p_, (300, 300))
        m_ = P.categorize_multilayer_image(r)
        m_ = P.erode_image(m_, 2)
        l_ = P.label_multilayer_image(m_)
        l_ = P.dilate_image(l_, 2)
        P.build_score(l_, r)
    return (time.perf_counter() - t0) / len(imgs) * 1e3


cpu_ms = cpu_postproc_ms_per_image(probs_syn)
out = {"workload": "UNetResNet-%d eval forward + softmax, batch %d @%dx%d, then mask post-processing to 300x300 (%s)" % (enc, b, s, s, mode),
       "forward_ms": round(t_fwd, 3), "postproc_ms": round(t_pp, 3), "postproc_share_of_step": round(t_pp / (t_fwd + t_pp), 4),
       "postproc_config": "erode 2 + dilate 2 (REPRODUCE_RESULTS.md evaluation setting)",
       "postproc_eager_launch_ms": round(t_pp_eager, 3), "postproc_default_eager_launch_ms": round(t_pp_default_eager, 3),
       "postproc_default_ms": round(t_pp_default, 3),
       "postproc_default_share_of_step": round(t_pp_default / (t_fwd + t_pp_default), 4),
       "postproc_default_config": "erode 0 / dilate 0 (neptune.yaml defaults)",
       "tiles_per_s_inference_plus_postproc": round(b / ((t_fwd + t_pp) * 1e-3), 1),
       "dense_crf_5iter_ms": round(t_crf, 3), "watershed_ms": round(t_ws, 3),
       "stages_ms": {k: round(v, 3) for k, v in stages.items()},
       "cpu_postproc_ms_per_image": round(cpu_ms, 2), "cpu_postproc_what": "oracle restatement of the reference's serial "
       "per-image chain (erode 2 + dilate 2) on one host core, 8 images",
       "gpu_postproc_ms_per_image": round(t_pp / b, 4),
       "postproc_algorithmic_MB": round(b * 1.44, 1), "postproc_GBps_vs_1.44MB_per_image": round(b * 1.44e-3 / (t_pp * 1e-3), 1)}
print(json.dumps(out))
