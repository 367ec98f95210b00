"""The library baseline of bench.py (baseline/torch_cudnn_unet.py: the reference's network and loss in stock PyTorch
modules) must be THE SAME FUNCTION as the pinned oracle: same state_dict keys, bit-identical fp32 logits and loss on the
CPU.  (On the GPU it runs under bf16 autocast / cuDNN; that is the thing being timed, not checked.)"""
import torch

from baseline.torch_cudnn_unet import UNetResNet, mixed_loss
from oracle import synthetic
from oracle import unet_oracle as O


def test_library_baseline_equals_oracle_on_cpu():
    sd = O.make_reference_like_state_dict(34, seed=3)
    net = UNetResNet(34)
    res = net.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    x, t = synthetic.train_batch(2, 64, seed=1, n_rect=5)
    X, T = torch.from_numpy(x), torch.from_numpy(t)
    net.train()
    with torch.no_grad():
        a = net(X)
        b = O.UNetOracle({k: v.clone() for k, v in sd.items()}, 34, update_running_stats=False).forward(X, training=True)
    assert torch.equal(a, b)
    assert float(mixed_loss(a, T)) == float(O.mixed_loss(b, T))
