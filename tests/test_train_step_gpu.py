"""The fused train step (PyTorchUNetWeighted._fit_loop on the B200 path) against the CPU oracle restatement of the
reference's Model._fit_loop, and the transformer surface (transform -> softmax probabilities)."""
import numpy as np
import pytest
import torch

from oracle import synthetic
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def test_fit_loop_tracks_oracle_for_several_steps(mcb, cuda):
    import bench
    from mcb200.models import PyTorchUNetWeighted
    sd = O.make_reference_like_state_dict(34, seed=99)
    model = PyTorchUNetWeighted(**bench.unet_config("ResNet34"))
    model.model.load_state_dict(sd)
    x, t = synthetic.train_batch(4, 128, seed=3, n_rect=10)
    X, T = torch.from_numpy(x), torch.from_numpy(t)
    sd_o = {k: v.clone() for k, v in sd.items()}
    opt = O.AdamOracle(lr=5e-4, weight_decay=1e-4)
    losses, ref_losses = [], []
    for i in range(4):
        out = model._fit_loop([X, T])
        assert set(out) == {"sum"}
        losses.append(float(out["sum"].data.cpu()))
        ref_losses.append(float(O.train_step(sd_o, 34, X, T, opt, imsize=(256, 256))[0]))
    # step 1 is the same weights: tight; later steps diverge slowly (bf16 gradients through Adam's normalisation)
    assert abs(losses[0] - ref_losses[0]) < 1e-3 * abs(ref_losses[0]), (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 0.05 * abs(b), (losses, ref_losses)
    assert losses[-1] < losses[0]
    # parameters moved like the oracle's: decoder tail tensors (accurate gradients) must agree closely
    got = model.model.state_dict()
    for k in ("final.weight", "final.bias", "dec0.conv.bias"):
        d0 = (sd_o[k] - sd[k]).norm()
        assert (got[k].cpu() - sd_o[k]).norm() < 0.2 * d0 + 1e-6, k
    # BN running statistics follow nn.BatchNorm2d's update
    for k in ("encoder.bn1.running_mean", "encoder.bn1.running_var"):
        assert torch.allclose(got[k].cpu(), sd_o[k], rtol=2e-2, atol=1e-3), k


def test_plain_ce_model_and_generic_autograd_path(mcb, cuda):
    import bench
    from mcb200.models import PyTorchUNet
    model = PyTorchUNet(**bench.unet_config("ResNet34"))
    x, t = synthetic.train_batch(2, 64, seed=8, n_rect=5)
    X, T = torch.from_numpy(x), torch.from_numpy(t[:, :1].copy())
    l0 = float(model._fit_loop([X, T])["sum"])
    ref = float(O.plain_ce_loss(O.UNetOracle(O.strip_module_prefix({k: v.cpu() for k, v in model.model.state_dict().items()}), 34,
                                             update_running_stats=False).forward(X, training=True), T))
    for _ in range(3):
        l1 = float(model._fit_loop([X, T])["sum"])
    assert l1 < l0
    # generic path: a user-supplied torch loss goes through the autograd bridge + torch.optim.Adam
    model2 = PyTorchUNet(**bench.unet_config("ResNet34"))
    model2._fused_loss = None
    model2.loss_function = [("multichannel_map", lambda out, tgt: torch.nn.functional.cross_entropy(out, tgt.squeeze(1).long()), 1.0)]
    a = float(model2._fit_loop([X, T])["sum"])
    for _ in range(3):
        b = float(model2._fit_loop([X, T])["sum"])
    assert b < a
    assert abs(a - l0) < 0.05 * abs(l0) + 0.05  # same init distribution, same data: same ballpark


def test_transform_returns_softmax_probabilities(mcb, cuda):
    import bench
    from mcb200.models import PyTorchUNet, PyTorchUNetStream
    sd = O.make_reference_like_state_dict(34, seed=5)
    model = PyTorchUNet(**bench.unet_config("ResNet34"))
    model.model.load_state_dict(sd)
    x, _ = synthetic.train_batch(5, 64, seed=1, n_rect=4)
    batches = [torch.from_numpy(x[:2]), [torch.from_numpy(x[2:4])], torch.from_numpy(x[4:])]
    out = model.transform((batches, len(batches)))
    p = out["multichannel_map_prediction"]
    assert isinstance(p, np.ndarray) and p.shape == (5, 2, 64, 64) and p.dtype == np.float32
    assert np.allclose(p.sum(1), 1.0, atol=1e-6)
    ref = torch.softmax(O.UNetOracle({k: v.clone() for k, v in sd.items()}, 34).forward(torch.from_numpy(x)), 1).numpy()
    assert np.abs(p - ref).max() < 1e-3
    stream = PyTorchUNetStream(**bench.unet_config("ResNet34"))
    stream.model.load_state_dict(sd)
    gen = stream.transform((batches, len(batches)))["multichannel_map_prediction"]
    got = np.stack(list(gen))
    assert np.abs(got - ref).max() < 1e-3
