"""__graft_entry__.smoke(): one tiny train step, one inference + post-processing call on cuda:0, checked against the
CPU oracle (the oracle is imported here as the checker only)."""
import os
import sys

import numpy as np
import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench
    from oracle import post_oracle as P, synthetic, unet_oracle as O
    from . import postprocessing as G
    from .models import PyTorchUNetWeighted

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    torch.cuda.set_device(0)
    sd = O.make_reference_like_state_dict(34, seed=1234)
    model = PyTorchUNetWeighted(**bench.unet_config("ResNet34"))
    model.model.load_state_dict(sd)
    x, t = synthetic.train_batch(2, 64, seed=1234, n_rect=6)
    X, T = torch.from_numpy(x), torch.from_numpy(t)
    # inference: logits -> softmax, against the fp32 oracle
    probs = model.transform(([X], 1))["multichannel_map_prediction"]
    ref = torch.softmax(O.UNetOracle({k: v.clone() for k, v in sd.items()}, 34).forward(X), 1).numpy()
    err = float(np.abs(probs - ref).max())
    assert err < 1e-3, "inference parity %g" % err
    # one fused train step against the oracle restatement of Model._fit_loop
    loss = float(model._fit_loop([X, T])["sum"])
    sd_o = {k: v.clone() for k, v in sd.items()}
    ref_loss = float(O.train_step(sd_o, 34, X, T, O.AdamOracle(5e-4, 1e-4), imsize=(256, 256))[0])
    assert abs(loss - ref_loss) < 1e-3 * abs(ref_loss), (loss, ref_loss)
    # post-processing chain, bit-exact labels
    pm = synthetic.probability_maps(2, 64, seed=7, n_rect=10)
    out = G.MaskPostprocessor((75, 75), "resize", 0, 2).transform(pm)["y_pred"]
    for p, (labels, scores) in zip(pm, out):
        r = P.resize_image(p, (75, 75))
        want = P.dilate_image(P.label_multilayer_image(P.categorize_multilayer_image(r)), 2)
        assert np.array_equal(labels, want), "label parity"
    torch.cuda.synchronize()
    print("smoke ok: inference max-abs %.2e, train loss %.6f (oracle %.6f), labels bit-exact" % (err, loss, ref_loss))
