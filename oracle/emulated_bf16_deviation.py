"""TEST INFRASTRUCTURE — how far does bf16 STORAGE alone move the training-mode logits and gradients away from the fp32
reference?  Runs the CPU oracle with the CUDA path's rounding points emulated (oracle.unet_oracle.UNetOracle(
emulate_bf16=True): bf16 conv operands, bf16-stored activations and activation gradients, fp32 accumulation) on the
conditioned checkpoints of tests/golden/unet_conditioned.npz and writes tests/golden/emulated_bf16_deviation.json:
per case and gradient tensor the relative L2 deviation and cosine against the reference's fp32 gradients.
tests/test_unet_configs_gpu.py bounds the CUDA path's own deviation by these numbers (+ a margin): the GPU may not be
further from the reference than the storage format itself puts a bit-faithful CPU evaluation.
    python -m oracle.emulated_bf16_deviation          (CPU, ~5 minutes)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synthetic, unet_oracle as O  # noqa: E402
from oracle.make_golden_cases import CONFIG_CASES, GRAD_HEAD, LOGIT_STRIDE  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "unet_conditioned.npz"))
    out = {}
    for tag, enc, depth, n, s in CONFIG_CASES:
        x, t = synthetic.train_batch(n, s, seed=1234)
        X, T = torch.from_numpy(x), torch.from_numpy(t)
        sd = O.conditioned_state_dict(depth, X, seed=1234)
        names = [k[len("grad_%s_" % tag):] for k in g.files if k.startswith("grad_%s_" % tag)]
        leaves = {k: sd[k].clone().requires_grad_(True) for k in O.trainable_keys(sd)}
        work = dict(sd)
        work.update(leaves)
        logits = O.UNetOracle(work, depth, update_running_stats=False, emulate_bf16=True).forward(X, training=True)
        loss = O.mixed_loss(logits, T, imsize=(256, 256))
        grads = torch.autograd.grad(loss, [leaves[k] for k in names])
        rec = {"logits_max_abs": float((logits.detach()[:, :, ::LOGIT_STRIDE, ::LOGIT_STRIDE] -
                                        torch.from_numpy(g["train_logits_" + tag])).abs().max()), "grads": {}}
        for k, gr in zip(names, grads):
            r = torch.from_numpy(g["grad_%s_%s" % (tag, k)]).double()
            a = gr.detach().reshape(-1)[:GRAD_HEAD].double()
            rec["grads"][k] = {"rel": float((a - r).norm() / r.norm()), "cos": float((a * r).sum() / (a.norm() * r.norm()))}
        out[tag] = rec
        print(tag, json.dumps(rec), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "emulated_bf16_deviation.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
