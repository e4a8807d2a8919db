"""Shared pieces of the two sampler modules (EMA weight swap, grid output, batch sharding)."""
import math
import os
from contextlib import contextmanager

import torch
import torchvision.utils as vutils

from .. import dist as _dist
from ..model.difussion_class import to_image01


@contextmanager
def sampling_weights(model, ema):
    """eval() + optional EMA swap, restored afterwards (ddpm_inference.py:22-28,54-56)."""
    was_training = model.training
    model.eval()
    backup = None
    if ema is not None:
        backup = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ema.copy_to(model)
    try:
        yield
    finally:
        if backup is not None:
            model.load_state_dict(backup)
        model.train(was_training)


def initial_noise(n, img_size, device, seed, shard):
    """x_T ~ N(0, I).  With `shard=True` under torch.distributed every rank draws the FULL batch
    from the same seed and keeps its own rows, so the union over ranks is bit-identical to a
    single-GPU run; no communication."""
    if seed is not None:
        torch.manual_seed(seed)
    x = torch.randn(n, 3, img_size, img_size, device=device)
    if shard:
        lo, hi = _dist.shard_range(n)
        x = x[lo:hi].contiguous()
    return x


def save_grid(x01, nrow, out_path, pad=2):
    grid = vutils.make_grid(x01, nrow=nrow, padding=pad)
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    vutils.save_image(grid, out_path)
    return grid


def save_each(x01, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for i in range(x01.shape[0]):
        vutils.save_image(x01[i], os.path.join(out_dir, f"img_{i:03d}.png"))


__all__ = ["sampling_weights", "initial_noise", "save_grid", "save_each", "to_image01", "math"]
