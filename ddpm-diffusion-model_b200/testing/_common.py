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


class StepGraph:
    """One sampler transition x_t -> x_{t-1} (UNet forward + the fused DDPM / DDIM kernel, ~150 launches)
    captured once as a CUDA graph and replayed per step: the reference's loop is launch-bound at small batch
    (two host syncs and ~120 ATen calls per step on top of the model).

    The random draw stays OUTSIDE the graph -- `torch.randn_like` on the default generator, in the same order
    as the reference (difussion_class.py:186,213) -- so seeds reproduce bit-identical noise.
    `step(x, t, t_prev, noise) -> x_next` must be a pure function of its tensor arguments."""

    def __init__(self, step, x: torch.Tensor, with_prev: bool):
        B = x.shape[0]
        dev = x.device
        self.x = x.clone()
        self.t = torch.zeros(B, dtype=torch.long, device=dev)
        self.tp = torch.zeros(B, dtype=torch.long, device=dev) if with_prev else None
        self.noise = torch.zeros_like(x)
        self._step = step
        # eager warm-up on a side stream: builds TMA-free host state (packed weights, schedule handle, pooled
        # activation buffers) so that nothing is allocated or packed inside the capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._call()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._call()

    def _call(self):
        return self._step(self.x, self.t, self.tp, self.noise)

    def run(self, cur: int, prev, draw_noise: bool = True) -> None:
        """Advance the internal state by one transition (state lives in `self.x`)."""
        self.t.fill_(cur)
        if self.tp is not None:
            self.tp.fill_(prev)
        if draw_noise:
            self.noise.copy_(torch.randn_like(self.noise))    # same call (and RNG stream) as difussion_class.py:186,213
        self.graph.replay()
        self.x.copy_(self.out)


def graphs_enabled() -> bool:
    """Opt-in (DDPM_B200_GRAPHS=1).  Measured on B200 (tools/graphdiag.py): a replayed step costs the same as the
    eagerly launched one at every batch size (B=8: 1.60 vs 1.75 ms, B=256: 8.45 vs 8.53 ms) -- the step is bound
    by GPU-side per-kernel latency, not by host launches -- so graphs are not the default."""
    return os.environ.get("DDPM_B200_GRAPHS", "0") == "1"


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
