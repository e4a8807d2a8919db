"""Drop-in for src/testing/ddpm_inference.py: T-step ancestral sampling.  Each step is one UNet
forward on libddpm_b200 plus ONE fused p_sample_step kernel; under torch.distributed the batch can
be sharded across ranks with no communication (`shard=True`)."""
import math

import torch

from ._common import StepGraph, graphs_enabled, initial_noise, sampling_weights, save_each, save_grid, to_image01


@torch.no_grad()
def ddpm_infer_sample(model, diffusion, n: int = 36, img_size: int = 64, device: str = "cuda", *, ema=None,
                      out_path: str = "samples_ddpm.png", save_individual: bool = False,
                      out_dir: str = "samples_individual", seed: int | None = 1234, shard: bool = False):
    """ddpm_inference.py:6-58.  Returns the image grid (C,H,W) in [0,1]."""
    with sampling_weights(model, ema):
        x = initial_noise(n, img_size, device, seed, shard)
        B = x.shape[0]
        if graphs_enabled() and diffusion.T > 3:
            sg = StepGraph(lambda x_, t_, tp_, z_: diffusion.p_sample_step(model, x_, t_, noise=z_), x, False)
            for i in reversed(range(diffusion.T)):
                sg.run(i, None)
            x = sg.x
        else:
            for i in reversed(range(diffusion.T)):
                t = torch.full((B,), i, device=x.device, dtype=torch.long)
                x = diffusion.p_sample_step(model, x, t)
        x = to_image01(x)
        grid = save_grid(x, int(math.sqrt(n)), out_path)
        print(f"[INFER] Grid guardado en: {out_path}")
        if save_individual:
            save_each(x, out_dir)
            print(f"[INFER] {B} imágenes individuales guardadas en: {out_dir}")
    return grid


@torch.no_grad()
def render_denoise_strip(model, diffusion, *, img_size: int = 64, device: str = "cuda", ema=None,
                         seed: int | None = 1234, out_path: str = "denoise_strip.png",
                         capture_steps: list[int] | None = None, pad: int = 2):
    """ddpm_inference.py:62-119: one sample, snapshots at `capture_steps` (default ~20 evenly spaced)."""
    T = diffusion.T
    if capture_steps is None:
        capture_steps = [int(v) for v in torch.linspace(T - 1, 0, 20).round().tolist()]
    wanted = set(capture_steps)
    frames = []
    with sampling_weights(model, ema):
        x = initial_noise(1, img_size, device, seed, False)
        for i in range(T - 1, -1, -1):
            t = torch.full((1,), i, device=x.device, dtype=torch.long)
            x = diffusion.p_sample_step(model, x, t)
            if i in wanted:
                frames.append(to_image01(x)[0])           # stays on the GPU; one D2H at the end
        grid = save_grid(torch.stack(frames, 0), len(frames), out_path, pad)
        print(f"[DENOISE] strip 1×{len(frames)} guardado → {out_path}")
    return grid
