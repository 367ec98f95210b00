"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, through
oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container only:  python -m oracle.make_golden
The fixtures pin the CPU oracle restatements (oracle/post_oracle.py, oracle/unet_oracle.py); the GPU tests then compare
the CUDA path with the oracle and with these fixtures."""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_postproc(pp, ut):
    probs = synthetic.probability_maps(3, 64, seed=1234, n_rect=12)
    rec = {"probs": probs}
    for i, p in enumerate(probs):
        r = pp.resize_image(p, (75, 75))
        c = pp.categorize_multilayer_image(r)
        l = pp.label_multilayer_image(c)
        d = pp.dilate_image(l, 2)
        d3 = pp.dilate_image(l, 3)
        e = pp.erode_image(c[1], 2)          # 2-D branch (the multi-layer branch of the reference is broken)
        e3 = pp.erode_image(c[1], 3)
        _, s = pp.build_score(d, r)
        crop = pp.crop_image_center_per_class(p, 56, 56)
        rec.update({"resize_%d" % i: r, "cat_%d" % i: c, "label_%d" % i: l, "dilate2_%d" % i: d, "dilate3_%d" % i: d3,
                    "erode2_%d" % i: e, "erode3_%d" % i: e3, "crop_%d" % i: crop,
                    "score0_%d" % i: np.array([float(v) for v in s[0]]),
                    "score1_%d" % i: np.array([float(v) for v in s[1]])})
    logits = np.random.RandomState(5).randn(2, 2, 16, 16).astype(np.float32) * 3
    rec["softmax_in"] = logits
    rec["softmax_out"] = ut.softmax(logits, axis=1)
    m = np.array([[0, 0, 1, 1], [1, 0, 0, 0], [1, 1, 1, 0], [0, 0, 1, 0]])
    rec["docstring_mask"] = m
    rec["docstring_labels"] = pp.label_multiclass_image(m)
    np.savez_compressed(os.path.join(OUT, "postproc.npz"), **rec)
    print("postproc.npz", len(rec), "arrays")


def golden_unet(um, mo):
    rec = {}
    for depth, enc in ((34, "ResNet34"),):
        cfg = ref_shim.reference_unet_config(enc)
        torch.manual_seed(1234)
        model = mo.PyTorchUNetWeighted(**cfg)   # builds UNetResNet(depth) + Adam(lr 5e-4, wd 1e-4) like the reference
        x, t = synthetic.train_batch(2, 64, seed=1234, n_rect=6)
        X, T = torch.from_numpy(x), torch.from_numpy(t)
        model.model.eval()
        with torch.no_grad():
            rec["eval_logits_%d" % depth] = model.model(X).numpy()
        model.model.train()
        # loss + gradients of the first step (without the update)
        out = model.model(X)
        name, loss_fn, weight = model.loss_function[0]
        loss = loss_fn(out, T) * weight
        loss.backward()
        rec["train_logits_%d" % depth] = out.detach().numpy()
        rec["loss_%d" % depth] = np.array(float(loss))
        sd = dict(model.model.named_parameters())
        for k in ("final.weight", "final.bias", "dec0.conv.weight", "dec1.block.1.weight", "center.block.0.conv.bias",
                  "encoder.layer1.0.conv1.weight", "encoder.bn1.weight", "encoder.conv1.weight"):
            rec["grad_%d_%s" % (depth, k)] = sd[k].grad.detach().numpy().copy()
        # undo the forward's running-stat update, then run the reference's own train step
        torch.manual_seed(1234)
        model = mo.PyTorchUNetWeighted(**cfg)
        res = model._fit_loop([X, T])
        rec["fit_loss_%d" % depth] = np.array(float(res["sum"]))
        sd2 = model.model.state_dict()
        for k in ("final.weight", "dec0.conv.bias", "encoder.bn1.running_mean", "encoder.bn1.running_var",
                  "encoder.layer4.2.bn2.weight", "center.block.1.bias"):
            rec["step_%d_%s" % (depth, k)] = sd2[k].numpy().copy()
        res2 = model._fit_loop([X, T])
        rec["fit_loss2_%d" % depth] = np.array(float(res2["sum"]))
    rec["x"], rec["t"] = x, t
    np.savez_compressed(os.path.join(OUT, "unet.npz"), **rec)
    print("unet.npz", {k: v.shape for k, v in rec.items() if v.ndim == 0 or "logits" in k})


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    um, mo, pp, ut = ref_shim.reference_modules()
    golden_postproc(pp, ut)
    golden_unet(um, mo)
