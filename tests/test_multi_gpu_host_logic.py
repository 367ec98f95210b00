"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: the loss is defined on the GLOBAL batch
(src/steps/pytorch/models.py:92-104 computes it on the gathered outputs), so each rank reduces its four partial sums,
all-reduces them, and forms loss + gradient locally; summed over ranks that equals the single-process result."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synthetic
from oracle import unet_oracle as O


def _partials(logits, target):
    """the four sums csrc/loss.cu's phase 1 produces, in torch (float64)"""
    w = O.loss_weights(target, 50.0, 10.0, (256, 256)).double()
    t = target[:, 0].double()
    z = logits.double()
    p1 = torch.softmax(z, 1)[:, 1]
    ce = torch.logsumexp(z, 1) - torch.where(t > 0.5, z[:, 1], z[:, 0])
    return torch.stack([(p1 * t).sum(), p1.sum(), t.sum(), (w * ce).sum()])


def _loss_from_sums(s, global_pixels, dice_w=0.2, ce_w=1.0, smooth=1.0, eps=1e-7):
    I, P, T, S = s
    return dice_w * (1 - (2 * I + smooth) / (P + T + smooth + eps)) + ce_w * S / global_pixels


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    logits_all = torch.randn(4, 2, 32, 32)
    _, t = synthetic.train_batch(4, 32, seed=9, n_rect=4)
    target_all = torch.from_numpy(t)
    lo, hi = rank * 2, rank * 2 + 2
    sums = _partials(logits_all[lo:hi], target_all[lo:hi])
    dist.all_reduce(sums)
    loss = _loss_from_sums(sums, 4 * 32 * 32)
    out[rank] = float(loss)
    dist.destroy_process_group()


def test_global_loss_from_all_reduced_partial_sums():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    logits_all = torch.randn(4, 2, 32, 32)
    _, t = synthetic.train_batch(4, 32, seed=9, n_rect=4)
    ref = float(O.mixed_loss(logits_all, torch.from_numpy(t), imsize=(256, 256)))
    assert abs(out[0] - out[1]) < 1e-12
    assert abs(out[0] - ref) < 1e-5 * abs(ref)


def test_rank_sharded_synthetic_batches_are_disjoint_and_deterministic():
    a0, _ = synthetic.train_batch(2, 32, seed=1234 + 0)
    a1, _ = synthetic.train_batch(2, 32, seed=1234 + 1)
    b0, _ = synthetic.train_batch(2, 32, seed=1234 + 0)
    assert (a0 == b0).all() and not (a0 == a1).all()
