"""Diagnostic: UNetResNet (B200 path) vs the CPU fp32 oracle — activations per stage, logits, gradients."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcb200
from mcb200.unet_models import UNetResNet
from oracle import unet_oracle as O

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 34
N, S = (int(sys.argv[2]) if len(sys.argv) > 2 else 2), (int(sys.argv[3]) if len(sys.argv) > 3 else 64)
dev = torch.device("cuda:0")
sd = O.make_reference_like_state_dict(depth, seed=1234)
net = UNetResNet(depth, 2, 32, 0.0, False, True)
missing = net.load_state_dict(sd, strict=True)
net = net.cuda()
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 3, S, S, generator=g)
target = torch.zeros(N, 3, S, S)
target[:, 0, S // 4:S // 2, S // 4:S // 2] = 1
target[:, 1] = torch.randint(0, 20, (N, S, S), generator=g).float() * (1 - target[:, 0])
target[:, 2] = 1 + target[:, 0] * 15


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    a, b = a.detach(), b.detach()
    return float((a - b).norm() / (b.norm() + 1e-12)), float((a - b).abs().max()), float(b.abs().max())


# ---- eval forward
net.eval()
with torch.no_grad():
    y = net(x.to(dev))
torch.cuda.synchronize()
sd_o = {k: v.clone() for k, v in sd.items()}
ref, inter = O.UNetOracle(sd_o, depth).forward(x, training=False, return_intermediates=True)
pl = net.plan(N, S, S, False)
for k, t in pl.named.items():
    print("eval %-7s rel %.4f maxerr %.4f refmax %.3f" % ((k,) + rel(t.permute(0, 3, 1, 2), inter[k])))
print("eval logits rel %.4f maxerr %.4f refmax %.3f" % rel(y, ref), flush=True)

# ---- train forward + backward
net.train()
sd_o = {k: v.clone() for k, v in sd.items()}
xg = x.to(dev)
t0 = time.time()
logits = net(xg)
loss = O.mixed_loss(logits, target.to(dev), imsize=(S, S))
loss.backward()
torch.cuda.synchronize()
print("first train step (eager + capture) %.2fs" % (time.time() - t0))
EMU = os.environ.get("EMU", "1") == "1"
oracle = O.UNetOracle(sd_o, depth, emulate_bf16=EMU)
print("oracle emulate_bf16 =", EMU)
keys = O.trainable_keys(O.strip_module_prefix(sd_o))
leaves = {k: sd_o[k].clone().requires_grad_(True) for k in keys}
work = dict(sd_o); work.update(leaves)
oracle.sd = work
ref_logits, inter = oracle.forward(x, training=True, return_intermediates=True)
ref_loss = O.mixed_loss(ref_logits, target, imsize=(S, S))
grads = torch.autograd.grad(ref_loss, [leaves[k] for k in keys], allow_unused=True)
pl = net.plan(N, S, S, True)
for k, t in pl.named.items():
    print("train %-7s rel %.4f maxerr %.4f refmax %.3f" % ((k,) + rel(t.permute(0, 3, 1, 2), inter[k])))
print("train logits rel %.4f maxerr %.4f refmax %.3f | loss %.6f ref %.6f" % (rel(logits, ref_logits) + (float(loss), float(ref_loss))))
params = dict(net.named_parameters())
worst = []
for k, gref in zip(keys, grads):
    if gref is None:
        continue
    p = params[k]
    r = rel(p.grad, gref)
    cos = float(torch.nn.functional.cosine_similarity(p.grad.float().cpu().flatten(), gref.flatten(), dim=0))
    worst.append((r[0], cos, k, r[2]))
print("all gradients in network order:")
for r, cos, k, m in worst:
    print("  %-45s rel %.4f cos %.4f refmax %.3g" % (k, r, cos, m))
worst.sort(reverse=True)
print("gradient check: %d tensors; median rel %.4f" % (len(worst), worst[len(worst) // 2][0]))
for r, cos, k, m in worst[:12]:
    print("  worst %-45s rel %.4f cos %.4f refmax %.3g" % (k, r, cos, m))
for r, cos, k, m in worst[-3:]:
    print("  best  %-45s rel %.4f cos %.4f refmax %.3g" % (k, r, cos, m))
# running stats
rm = net.encoder.bn1.running_mean.cpu(); print("bn1 running_mean rel", rel(rm, work["encoder.bn1.running_mean"]))
# replay timing
for _ in range(3):
    logits = net(xg); O.mixed_loss(logits, target.to(dev), imsize=(S, S)).backward()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5):
    logits = net(xg); O.mixed_loss(logits, target.to(dev), imsize=(S, S)).backward()
torch.cuda.synchronize(); print("graph-replay train fwd+bwd: %.2f ms/iter; launches fwd %d bwd %d" % ((time.time() - t0) / 5 * 1e3, pl.launches_fwd, pl.launches_bwd))
