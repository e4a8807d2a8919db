"""Drop-in for src/testing/ddpim_inference.py: DDIM sampling.  The timestep schedule is built on
the host once (the reference reads two device scalars per step, ddpim_inference.py:77-78)."""
import math

import torch

from ._common import StepGraph, graphs_enabled, initial_noise, sampling_weights, save_each, save_grid, to_image01


def build_ddim_schedule(diffusion, steps: int, schedule_kind: str = "t_linear", schedule_idx=None) -> list:
    """ddpim_inference.py:40-70 -> descending list of ints ending in 0."""
    T = diffusion.T
    if schedule_idx is not None:
        s = sorted({int(v) for v in schedule_idx}, reverse=True)
    elif schedule_kind == "t_linear":
        pts = torch.linspace(T - 1, 0, steps).round().long()
        s = [int(v) for v in torch.unique_consecutive(pts)]
    elif schedule_kind == "alpha_bar_cosine":
        ab = diffusion.alphas_cumprod.detach().float().cpu()
        # same float32 linspace / subtraction as the reference so argmin ties resolve identically
        targets = 1.0 - torch.linspace(0.0, 1.0, steps)
        s = sorted({int((ab - z).abs().argmin()) for z in targets}, reverse=True)
    else:
        raise ValueError(f"schedule_kind desconocido: {schedule_kind}")
    if s[-1] != 0:
        s.append(0)
    return s


@torch.no_grad()
def ddim_infer_sample(model, diffusion, n: int = 36, img_size: int = 64, device: str = "cuda", *, ema=None,
                      out_path: str = "samples_ddim.png", save_individual: bool = False,
                      out_dir: str = "samples_individual", seed: int | None = 1234, steps: int = 50,
                      eta: float = 0.0, schedule_kind: str = "t_linear", schedule_idx: list[int] | None = None,
                      shard: bool = False):
    """ddpim_inference.py:7-104.  `steps` schedule points => steps-1 UNet evaluations."""
    with sampling_weights(model, ema):
        x = initial_noise(n, img_size, device, seed, shard)
        B = x.shape[0]
        sched = build_ddim_schedule(diffusion, steps, schedule_kind, schedule_idx)
        if graphs_enabled() and len(sched) > 3:
            sg = StepGraph(lambda x_, t_, tp_, z_: diffusion.p_sample_step_ddim(model, x_t=x_, t=t_, t_prev=tp_, eta=eta,
                                                                                clip_x0=True, noise=z_), x, True)
            for cur, prev in zip(sched[:-1], sched[1:]):
                sg.run(cur, prev)                      # noise is drawn every step, like the reference (App. C.5)
            x = sg.x
        else:
            for cur, prev in zip(sched[:-1], sched[1:]):
                t = torch.full((B,), cur, device=x.device, dtype=torch.long)
                tp = torch.full((B,), prev, device=x.device, dtype=torch.long)
                x = diffusion.p_sample_step_ddim(model, x_t=x, t=t, t_prev=tp, eta=eta, clip_x0=True, noise=None)
        x = to_image01(x)
        r = int(math.sqrt(n))
        grid = save_grid(x, r if r * r == n else math.ceil(math.sqrt(n)), out_path)
        print(f"[INFER-DDIM] Grid → {out_path}  (steps={len(sched) - 1}, eta={eta}, schedule={schedule_kind})")
        if save_individual:
            save_each(x, out_dir)
    return grid


@torch.no_grad()
def render_denoise_strip_ddim(model, diffusion, *, img_size: int = 64, device: str = "cuda", ema=None, seed=1234,
                              out_path: str = "denoise_strip_ddim.png", capture_steps=None, pad: int = 2,
                              steps: int = 50, eta: float = 0.0, schedule_kind: str = "linear", schedule_idx=None):
    """ddpim_inference.py:108-197: one DDIM trajectory, ~17 snapshots."""
    T = diffusion.T
    if schedule_idx is not None:
        sched = sorted({int(v) for v in schedule_idx}, reverse=True)
    elif schedule_kind == "cosine":
        u = torch.linspace(0, 1, steps)
        sched = sorted(set(torch.round((T - 1) * (1 - 0.5 * (1 - torch.cos(math.pi * u)))).long().tolist()), reverse=True)
    else:
        sched = sorted(set(torch.round(torch.linspace(T - 1, 0, steps)).long().tolist()), reverse=True)
    if capture_steps is None:
        k = min(17, len(sched))
        capture_steps = [sched[i] for i in torch.linspace(0, len(sched) - 1, k).round().long().tolist()]
    wanted = {int(v) for v in capture_steps}
    frames = []
    with sampling_weights(model, ema):
        x = initial_noise(1, img_size, torch.device(device), seed, False)
        for i, cur in enumerate(sched):
            prev = sched[i + 1] if i + 1 < len(sched) else 0
            t = torch.full((1,), cur, device=x.device, dtype=torch.long)
            tp = torch.full((1,), prev, device=x.device, dtype=torch.long)
            x = diffusion.p_sample_step_ddim(model, x_t=x, t=t, t_prev=tp, eta=eta, clip_x0=True, noise=None)
            if cur in wanted:
                frames.append(to_image01(x)[0])
        grid = save_grid(torch.stack(frames, 0), len(frames), out_path, pad)     # frames stay on the GPU; one D2H inside
        print(f"[DENOISE-DDIM] strip 1×{len(frames)} guardado → {out_path} (steps={len(sched)}, eta={eta})")
    return grid
