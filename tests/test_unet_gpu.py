"""End-to-end parity of the CUDA U-Net (mcb200.unet_models.UNetResNet through the C ABI) against
  (a) the golden logits produced by the UNMODIFIED reference (tests/golden/unet.npz)  — north-star bound 1e-3 max-abs,
  (b) the CPU oracle in bf16-storage emulation mode (same rounding points as the CUDA path)     — tight,
  (c) the fp32 oracle for gradients, by cosine similarity (bf16 storage through 30+ BatchNorm layers of a random-init
      net is noisy; (b) is the sharp check of the plan's logic)."""
import os

import numpy as np
import pytest
import torch

from oracle import synthetic
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGIT_TOL = 1e-3  # BASELINE.json north_star: "within 1e-3 max-abs on logits"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.fixture(scope="module")
def setup34(mcb, cuda):
    from mcb200.unet_models import UNetResNet
    sd = O.make_reference_like_state_dict(34, seed=1234)
    net = UNetResNet(34, 2, 32, 0.0, False, True)
    net.load_state_dict(sd, strict=True)
    net.cuda()
    return net, sd


def test_eval_logits_match_reference_golden(setup34, cuda):
    net, sd = setup34
    g = np.load(os.path.join(GOLD, "unet.npz"))
    net.load_state_dict(sd)
    net.eval()
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]).to(cuda))
    err = np.abs(y.cpu().numpy() - g["eval_logits_34"]).max()
    assert y.shape == (2, 2, 64, 64) and y.dtype == torch.float32
    assert err < LOGIT_TOL, err


def test_train_logits_and_loss_match_reference_golden(setup34, cuda):
    from mcb200 import models
    net, sd = setup34
    g = np.load(os.path.join(GOLD, "unet.npz"))
    net.load_state_dict(sd)
    net.train()
    X, T = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["t"]).to(cuda)
    logits = net(X)
    assert np.abs(logits.detach().cpu().numpy() - g["train_logits_34"]).max() < LOGIT_TOL
    loss = models.mixed_dice_cross_entropy_loss(logits, T, dice_weight=0.2, cross_entropy_weight=1.0, smooth=1,
                                                w0=50, sigma=10, imsize=(256, 256))
    assert abs(float(loss) - float(g["loss_34"])) < 2e-4 * abs(float(g["loss_34"]))
    loss.backward()
    params = dict(net.named_parameters())
    for name in g.files:
        if name.startswith("grad_34_"):
            k = name[len("grad_34_"):]
            ref = torch.from_numpy(g[name])
            got = params[k].grad.detach().cpu()
            assert got.shape == ref.shape and bool(torch.isfinite(got).all()), k
            # at 2x64x64 the deep layers see 2..32 samples per BatchNorm channel: their gradients are chaotic under any
            # bf16 storage; the decoder tail is well conditioned and must match tightly
            if k.startswith(("final", "dec0", "dec1")):
                assert rel(got, ref) < 2e-2, (k, rel(got, ref))


@pytest.mark.parametrize("depth,n,s", [(34, 4, 128), (101, 2, 64)])
def test_forward_backward_against_bf16_emulated_oracle(mcb, cuda, depth, n, s):
    from mcb200.unet_models import UNetResNet
    sd = O.make_reference_like_state_dict(depth, seed=4321)
    net = UNetResNet(depth, 2, 32, 0.0, False, True)
    net.load_state_dict(sd)
    net.cuda().train()
    x, t = synthetic.train_batch(n, s, seed=5, n_rect=8)
    X, T = torch.from_numpy(x), torch.from_numpy(t)
    logits = net(X.to(cuda))
    loss = O.mixed_loss(logits, T.to(cuda), imsize=(256, 256))
    loss.backward()
    # oracle with the same storage roundings
    sd_o = {k: v.clone() for k, v in sd.items()}
    keys = O.trainable_keys(sd_o)
    leaves = {k: sd_o[k].clone().requires_grad_(True) for k in keys}
    work = dict(sd_o)
    work.update(leaves)
    ref_logits, inter = O.UNetOracle(work, depth, emulate_bf16=True).forward(X, training=True, return_intermediates=True)
    ref_loss = O.mixed_loss(ref_logits, T, imsize=(256, 256))
    grads = dict(zip(keys, torch.autograd.grad(ref_loss, [leaves[k] for k in keys], allow_unused=True)))
    plan = net.plan(n, s, s, True)
    for name, tns in plan.named.items():
        assert rel(tns.permute(0, 3, 1, 2), inter[name]) < 3e-2, (name, rel(tns.permute(0, 3, 1, 2), inter[name]))
    assert float((logits.detach().cpu() - ref_logits.detach()).abs().max()) < LOGIT_TOL
    assert abs(float(loss) - float(ref_loss)) < 1e-3 * abs(float(ref_loss))
    params = dict(net.named_parameters())
    rels = []
    for k in keys:
        if grads[k] is None:
            continue
        r = rel(params[k].grad, grads[k])
        cos = float(torch.nn.functional.cosine_similarity(params[k].grad.detach().cpu().flatten(), grads[k].flatten(), dim=0))
        rels.append((r, cos, k))
    rels.sort(reverse=True)
    med = rels[len(rels) // 2][0]
    assert med < 0.1, (med, rels[:5])
    assert min(c for _, c, _ in rels) > 0.8, rels[:5]
    # running statistics updated like nn.BatchNorm2d
    assert rel(net.encoder.bn1.running_mean, work["encoder.bn1.running_mean"]) < 1e-2
    assert rel(net.encoder.bn1.running_var, work["encoder.bn1.running_var"]) < 1e-2


def test_state_dict_roundtrip_and_module_prefix(setup34, cuda, tmp_path):
    from mcb200.models import PyTorchUNet
    import bench
    net, sd = setup34
    net.load_state_dict(sd)
    out = net.state_dict()
    assert set(out) == set(sd)
    for k in sd:
        assert torch.equal(out[k].cpu(), sd[k]), k
    m = PyTorchUNet(**bench.unet_config("ResNet34"))
    m.model.load_state_dict(sd)
    path = str(tmp_path / "transformers" / "unet")
    m.save(path)
    saved = torch.load(path)
    assert all(k.startswith("module.") for k in saved)
    m2 = PyTorchUNet(**bench.unet_config("ResNet34"))
    m2.load(path)
    for k, v in m2.model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_rejects_bad_inputs(setup34, cuda):
    net, _ = setup34
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 300, 300, device=cuda))  # not a multiple of 64 (the reference fails in torch.cat)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))  # CPU tensor: no fallback
