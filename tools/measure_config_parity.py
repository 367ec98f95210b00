"""GPU box: deviation of the CUDA path from the reference goldens at the BASELINE.json configurations
(tests/golden/unet_configs.npz) -- prints the numbers the parity tests assert on."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcb200
from mcb200.unet_models import UNetResNet
from mcb200 import models
from oracle import synthetic, unet_oracle as O
from oracle.make_golden_cases import CONFIG_CASES, GRAD_HEAD, LOGIT_STRIDE

FAMILY = sys.argv[1] if len(sys.argv) > 1 else "configs"
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "unet_%s.npz" % FAMILY))
S3 = LOGIT_STRIDE if FAMILY == "conditioned" else 1
dev = torch.device("cuda:0")
print("family", FAMILY)
for tag, enc, depth, n, s in CONFIG_CASES:
    x0, _ = synthetic.train_batch(n, s, seed=1234)
    sd = (O.conditioned_state_dict(depth, torch.from_numpy(x0), seed=1234) if FAMILY == "conditioned"
          else O.make_reference_like_state_dict(depth, seed=1234))
    net = UNetResNet(depth, 2, 32, 0.0, False, True)
    net.load_state_dict(sd)
    net.cuda()
    x, t = synthetic.train_batch(n, s, seed=1234)
    X, T = torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev)
    net.eval()
    with torch.no_grad():
        y = net(X[:1]).cpu().numpy()[:, :, ::S3, ::S3]
    ref = g["eval_logits_" + tag]
    print(tag, "eval  max-abs %.3e  (ref range %.3f..%.3f, std %.3e)" % (np.abs(y - ref).max(), ref.min(), ref.max(), ref.std()))
    net.load_state_dict(sd)
    net.train()
    logits = net(X)
    ref = g["train_logits_" + tag]
    print(tag, "train max-abs %.3e  (ref range %.3f..%.3f, std %.3e)" % (np.abs(logits.detach().cpu().numpy()[:, :, ::S3, ::S3] - ref).max(), ref.min(), ref.max(), ref.std()))
    loss = models.mixed_dice_cross_entropy_loss(logits, T, dice_weight=0.2, cross_entropy_weight=1.0, smooth=1, w0=50, sigma=10, imsize=(256, 256))
    print(tag, "loss %.6f ref %.6f rel %.2e" % (float(loss), float(g["loss_" + tag]), abs(float(loss) - float(g["loss_" + tag])) / abs(float(g["loss_" + tag]))))
    loss.backward()
    params = dict(net.named_parameters())
    for name in g.files:
        if name.startswith("grad_%s_" % tag):
            k = name[len("grad_%s_" % tag):]
            r = torch.from_numpy(g[name])
            got = params[k].grad.detach().cpu().contiguous().reshape(-1)[:GRAD_HEAD]
            rel = float((got - r).norm() / (r.norm() + 1e-30))
            cos = float((got * r).sum() / (got.norm() * r.norm() + 1e-30))
            print("   grad %-40s rel %.3e cos %.5f |ref| %.3e" % (k, rel, cos, float(r.norm())))
    del net
    torch.cuda.empty_cache()
