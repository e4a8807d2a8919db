"""CPU oracle for the DDPM/DDIM hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, *functional* restatement (plain fp32 PyTorch on the CPU, no
nn.Module, weights taken from a ``state_dict``-shaped mapping) of the algorithm implemented by
pablo-reyes8/ddpm-diffusion-model.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``ddpm_diffusion_model_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  The reference ships no golden vectors; the oracle is
pinned by fixtures generated from the *live, unmodified reference* imported from
``/root/reference`` in the authoring container (``tools/make_golden.py`` -> ``tests/golden/*.pt``)
and checked by ``tests/test_oracle_golden.py``.

Each function cites the reference ``file:line`` (relative to /root/reference) that it follows.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# Schedules and the ten [T] tables        (src/model/difussion_utils.py:16-40,
#                                          src/model/difussion_class.py:36-68)
# --------------------------------------------------------------------------------------


def betas_linear(T: int, lo: float = 1e-4, hi: float = 2e-2) -> Tensor:
    """difussion_utils.py:16-20."""
    return torch.linspace(lo, hi, T, dtype=torch.float32)


def betas_cosine(T: int, s: float = 0.008) -> Tensor:
    """difussion_utils.py:22-40 (Nichol & Dhariwal alpha-bar, differenced)."""
    u = torch.arange(T + 1, dtype=torch.float32) / T
    ab = torch.cos((math.pi / 2.0) * ((u + s) / (1.0 + s))).clamp(min=1e-7) ** 2
    ab = ab / ab[0]
    return (1 - (ab[1:] / ab[:-1])).clamp(min=1e-8, max=0.999)


TABLE_NAMES = (
    "betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "alphas_cumprod_prev", "posterior_variance",
    "posterior_log_variance", "posterior_mean_coef1", "posterior_mean_coef2",
)


def make_tables(T: int = 1000, schedule: str = "linear", beta_min: float = 1e-4,
                beta_max: float = 2e-2, cosine_s: float = 0.008) -> Dict[str, Tensor]:
    """difussion_class.py:36-68.  Same torch ops in the same order => bit-identical tables."""
    if schedule == "linear":
        b = betas_linear(T, beta_min, beta_max)
    elif schedule == "cosine":
        b = betas_cosine(T, cosine_s)
    else:
        raise ValueError(f"schedule desconocido: {schedule}")
    a = 1.0 - b
    ab = torch.cumprod(a, dim=0)
    ab_prev = F.pad(ab[:-1], (1, 0), value=1.0)
    pv = b * (1.0 - ab_prev) / (1.0 - ab)
    return {
        "betas": b,
        "alphas": a,
        "alphas_cumprod": ab,
        "sqrt_alphas_cumprod": torch.sqrt(ab),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ab),
        "alphas_cumprod_prev": ab_prev,
        "posterior_variance": pv.clamp(min=1e-20),
        "posterior_log_variance": torch.log(pv.clamp(min=1e-20)),
        "posterior_mean_coef1": b * torch.sqrt(ab_prev) / (1.0 - ab),
        "posterior_mean_coef2": (1.0 - ab_prev) * torch.sqrt(a) / (1.0 - ab),
    }


def gather_coef(table: Tensor, t: Tensor, ndim: int = 4) -> Tensor:
    """difussion_utils.py:7-14: truncate to int64, clamp to [0,T-1], gather, view (B,1,1,1).

    (The reference clamps the caller's int64 ``t`` in place; the oracle does not mutate.)"""
    idx = t.long().clamp(0, table.shape[0] - 1)
    return table[idx].view((idx.shape[0],) + (1,) * (ndim - 1))


# --------------------------------------------------------------------------------------
# Diffusion elementwise math              (src/model/difussion_class.py:81-234)
# --------------------------------------------------------------------------------------


def q_sample(tb: Mapping[str, Tensor], x0: Tensor, t: Tensor, eps: Tensor) -> Tensor:
    """difussion_class.py:81-91."""
    return (gather_coef(tb["sqrt_alphas_cumprod"], t, x0.ndim) * x0
            + gather_coef(tb["sqrt_one_minus_alphas_cumprod"], t, x0.ndim) * eps)


def loss_simple(tb, eps_fn: Callable[[Tensor, Tensor], Tensor], x0: Tensor, t: Tensor,
                noise: Tensor, weight: Optional[Tensor] = None) -> Tensor:
    """difussion_class.py:95-116."""
    x_t = q_sample(tb, x0, t, noise)
    pred = eps_fn(x_t, t)
    per = (noise - pred).pow(2).mean(dim=(1, 2, 3))
    if weight is not None:
        per = per * weight
    return per.mean()


def predict_x0(tb, x_t: Tensor, eps: Tensor, t: Tensor, clamp_x0: bool = True,
               dynamic_threshold: Optional[float] = None) -> Tensor:
    """difussion_class.py:132-152."""
    sa = gather_coef(tb["sqrt_alphas_cumprod"], t, x_t.ndim)
    so = gather_coef(tb["sqrt_one_minus_alphas_cumprod"], t, x_t.ndim)
    x0 = (x_t - so * eps) / (sa + 1e-12)
    if dynamic_threshold is not None:
        amax = x0.abs().flatten(1).max(dim=1).values
        amax = torch.maximum(amax, torch.ones((), dtype=x0.dtype, device=x0.device))
        div = amax.clamp(min=dynamic_threshold).view(-1, *([1] * (x0.ndim - 1)))
        x0 = (x0 / div).clamp(-1, 1)
    elif clamp_x0:
        x0 = x0.clamp(-1, 1)
    return x0


def ddpm_step(tb, eps: Tensor, x_t: Tensor, t: Tensor, noise: Tensor, clamp_x0: bool = True,
              dynamic_threshold: Optional[float] = None, clip_x0: Optional[bool] = None) -> Tensor:
    """difussion_class.py:156-187 with the model output ``eps`` supplied."""
    if clip_x0 is None:
        clip_x0 = clamp_x0
    x0 = predict_x0(tb, x_t, eps, t, clamp_x0, dynamic_threshold)
    if clip_x0:
        x0 = x0.clamp(-1, 1)
    mean = (gather_coef(tb["posterior_mean_coef1"], t, x_t.ndim) * x0
            + gather_coef(tb["posterior_mean_coef2"], t, x_t.ndim) * x_t)
    logvar = gather_coef(tb["posterior_log_variance"], t, x_t.ndim)
    nz = (t > 0).float().view(-1, *([1] * (x_t.ndim - 1)))
    return mean + nz * torch.exp(0.5 * logvar) * noise


def ddim_step(tb, eps: Tensor, x_t: Tensor, t: Tensor, t_prev: Tensor, noise: Tensor,
              eta: float = 0.0, clamp_x0: bool = True, dynamic_threshold: Optional[float] = None,
              clip_x0: Optional[bool] = None) -> Tensor:
    """difussion_class.py:189-234 with the model output ``eps`` supplied."""
    if clip_x0 is None:
        clip_x0 = clamp_x0
    a_t = gather_coef(tb["alphas_cumprod"], t, x_t.ndim)
    a_p = gather_coef(tb["alphas_cumprod"], t_prev, x_t.ndim)
    x0 = predict_x0(tb, x_t, eps, t, clamp_x0, dynamic_threshold)
    if clip_x0:
        x0 = x0.clamp(-1, 1)
    direction = (x_t - torch.sqrt(a_t) * x0) / torch.sqrt(1.0 - a_t + 1e-12)
    sigma = eta * torch.sqrt((1.0 - a_p) / (1.0 - a_t + 1e-12)) * torch.sqrt(1.0 - a_t / (a_p + 1e-12))
    return (torch.sqrt(a_p) * x0
            + torch.sqrt(torch.clamp(1.0 - a_p - sigma ** 2, min=0.0)) * direction
            + sigma * noise)


# --------------------------------------------------------------------------------------
# UNet forward, functional, from a state_dict     (src/model/attention.py, unet_backbone.py)
# --------------------------------------------------------------------------------------


class UNetSpec:
    """Constructor arguments of UNetDenoiser (unet_backbone.py:78-88), as plain data."""

    def __init__(self, in_channels=3, base_channels=128, channel_mults=(1, 2, 2, 2),
                 num_res_blocks=2, attn_resolutions=frozenset({16, 8}), time_embed_dim=512,
                 dropout=0.0, num_heads=4, head_dim=64, img_resolution=64):
        self.in_channels = in_channels
        self.base_channels = base_channels
        self.channel_mults = tuple(channel_mults)
        self.num_res_blocks = num_res_blocks
        self.attn_resolutions = frozenset(attn_resolutions)
        self.time_embed_dim = time_embed_dim
        self.dropout = dropout
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.img_resolution = img_resolution

    def as_kwargs(self) -> dict:
        return dict(self.__dict__)


def sinusoidal(t: Tensor, dim: int) -> Tensor:
    """attention.py:13-22 (note the ``half-1`` denominator and zero pad for odd dim)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).to(t.device)
    ang = t.float()[:, None] * freq[None, :]
    out = torch.cat([ang.sin(), ang.cos()], dim=1)
    if dim % 2 == 1:
        out = F.pad(out, (0, 1))
    return out


def _gn(x: Tensor, sd, prefix: str) -> Tensor:
    """attention.py:38-39: GroupNorm(min(32,C) groups, eps 1e-6, affine)."""
    C = x.shape[1]
    return F.group_norm(x, min(32, C), sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-6)


def _conv(x: Tensor, sd, prefix: str, stride: int = 1, padding: int = 0) -> Tensor:
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)


def resblock(x: Tensor, temb: Tensor, sd, p: str) -> Tensor:
    """unet_backbone.py:37-44 (dropout off: eval / p=0)."""
    h = _conv(F.silu(_gn(x, sd, p + ".norm1")), sd, p + ".conv1", padding=1)
    tb = F.linear(F.silu(temb), sd[p + ".time_proj.1.weight"], sd[p + ".time_proj.1.bias"])
    h = h + tb[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + ".norm2")), sd, p + ".conv2", padding=1)
    if (p + ".skip.weight") in sd:
        x = _conv(x, sd, p + ".skip")
    return h + x


def attnblock(x: Tensor, sd, p: str, heads: int, head_dim: int) -> Tensor:
    """attention.py:56-74, with SDPA written out: softmax(q k^T / sqrt(d)) v over H*W tokens."""
    B, C, H, W = x.shape
    N = H * W
    qkv = _conv(_gn(x, sd, p + ".norm"), sd, p + ".qkv").reshape(B, 3, heads, head_dim, N)
    q, k, v = (qkv[:, i].transpose(-1, -2) for i in range(3))       # (B, heads, N, d)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(head_dim), dim=-1)
    o = (att @ v).transpose(-1, -2).reshape(B, heads * head_dim, H, W)
    return x + _conv(o, sd, p + ".proj")


def unet_layout(spec: UNetSpec):
    """Replays the constructor's bookkeeping (unet_backbone.py:102-163) and returns the list of
    per-level block kinds so the functional forward knows which state_dict prefixes exist."""
    downs, res, in_ch = [], spec.img_resolution, spec.base_channels
    skip_ch = []
    n_levels = len(spec.channel_mults)
    for li, mult in enumerate(spec.channel_mults):
        out_ch = spec.base_channels * mult
        blocks = []
        for _ in range(spec.num_res_blocks):
            blocks.append(("res", in_ch, out_ch))
            in_ch = out_ch
            if res in spec.attn_resolutions:
                blocks.append(("attn", in_ch, in_ch))
        skip_ch.append(in_ch)
        last = li == n_levels - 1
        downs.append({"blocks": blocks, "down": (not last), "ch": in_ch})
        if not last:
            res //= 2
    mid_attn = res in spec.attn_resolutions
    ups, cur = [], in_ch
    for li, mult in enumerate(reversed(spec.channel_mults)):
        out_ch = spec.base_channels * mult
        sk = list(reversed(skip_ch))[li]
        blocks = [("res", cur + sk, out_ch)] + [("res", out_ch, out_ch)] * spec.num_res_blocks
        ups.append({"blocks": blocks, "up": li != 0, "up_ch": cur})
        cur = out_ch
    return {"downs": downs, "mid_ch": in_ch, "mid_attn": mid_attn, "ups": ups}


def unet_forward(sd: Mapping[str, Tensor], spec: UNetSpec, x: Tensor, t: Tensor) -> Tensor:
    """unet_backbone.py:166-216."""
    lay = unet_layout(spec)
    e = sinusoidal(t, spec.time_embed_dim)
    e = F.linear(e, sd["time_mlp.net.0.weight"], sd["time_mlp.net.0.bias"])
    temb = F.linear(F.silu(e), sd["time_mlp.net.2.weight"], sd["time_mlp.net.2.bias"])

    cur = _conv(x, sd, "in_conv", padding=1)
    skips = []
    for li, lvl in enumerate(lay["downs"]):
        for bi, (kind, _, _) in enumerate(lvl["blocks"]):
            p = f"downs.{li}.blocks.{bi}"
            cur = resblock(cur, temb, sd, p) if kind == "res" else attnblock(cur, sd, p, spec.num_heads, spec.head_dim)
        skips.append(cur)
        if lvl["down"]:
            cur = _conv(cur, sd, f"downs.{li}.down.conv", stride=2, padding=1)
    cur = resblock(cur, temb, sd, "mid.0")
    if lay["mid_attn"]:
        cur = attnblock(cur, sd, "mid.1", spec.num_heads, spec.head_dim)
    cur = resblock(cur, temb, sd, "mid.2")
    for li, lvl in enumerate(lay["ups"]):
        if lvl["up"]:
            cur = F.interpolate(cur, scale_factor=2, mode="nearest")
            cur = _conv(cur, sd, f"ups.{li}.up.conv", padding=1)
        sk = skips.pop()
        if cur.shape[-2:] != sk.shape[-2:]:
            cur = F.interpolate(cur, size=sk.shape[-2:], mode="nearest")
        cur = torch.cat([cur, sk], dim=1)
        for bi in range(len(lvl["blocks"])):
            cur = resblock(cur, temb, sd, f"ups.{li}.blocks.{bi}")
    return _conv(F.silu(_gn(cur, sd, "out_norm")), sd, "out_conv", padding=1)


def unet_loss_and_grads(sd: Mapping[str, Tensor], spec: UNetSpec, tb, x0, t, noise,
                        loss_scale: float = 1.0) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """loss_simple + backward through the functional UNet via CPU autograd.
    Returns (loss, eps_pred, {name: grad})."""
    leaves = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}
    box = {}

    def fn(x_t, tt):
        box["eps"] = unet_forward(leaves, spec, x_t, tt)
        return box["eps"]

    loss = loss_simple(tb, fn, x0, t, noise)
    (loss * loss_scale).backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.detach(), box["eps"].detach(), grads


# --------------------------------------------------------------------------------------
# Optimiser-side parameter pass   (train_one_epoch.py:86-115, ema.py:15-23, torch AdamW)
# --------------------------------------------------------------------------------------


def unscale_and_clip(grads: Sequence[Tensor], inv_scale: float, max_norm: Optional[float]):
    """GradScaler.unscale_ (x 1/scale + found-inf) then clip_grad_norm_ (train_one_epoch.py:101-105):
    total = ||g||_2 over all tensors, coef = max_norm/(total+1e-6) clamped to 1."""
    g = [x * inv_scale for x in grads]
    found_inf = any((not torch.isfinite(x).all()) for x in g)
    total = torch.sqrt(sum((x.double() ** 2).sum() for x in g)).float()
    if max_norm is not None:
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        g = [x * coef for x in g]
    return g, float(total), found_inf


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, wd: float = 0.0):
    """torch.optim.AdamW single-tensor rule (decoupled decay, bias-corrected, eps outside sqrt
    of the bias-corrected second moment).  ``step`` is the 1-based step count."""
    p = p * (1.0 - lr * wd)
    m = m + (g - m) * (1.0 - beta1)                 # torch: exp_avg.lerp_(grad, 1-beta1)
    v = v * beta2 + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def ema_update(shadow: Tensor, p: Tensor, decay: float) -> Tensor:
    """ema.py:22: shadow.mul_(decay).add_(p, alpha=1-decay)."""
    return shadow * decay + p * (1.0 - decay)


def train_step(sd, spec: UNetSpec, tb, x0, t, noise, opt_state, ema_shadow, *, lr, step,
               betas=(0.9, 0.999), eps=1e-8, wd=0.0, grad_clip=1.0, ema_decay=0.999,
               loss_scale=1.0):
    """One optimiser step of train_one_epoch.py:61-121 in fp32 (no autocast, dropout 0):
    loss -> backward -> unscale -> clip -> AdamW -> EMA.  Mutates nothing; returns new states."""
    loss, eps_pred, grads = unet_loss_and_grads(sd, spec, tb, x0, t, noise, loss_scale)
    names = list(sd.keys())
    g, gnorm, found_inf = unscale_and_clip([grads[n] for n in names], 1.0 / loss_scale, grad_clip)
    new_sd, new_opt, new_ema = {}, {}, {}
    for n, gi in zip(names, g):
        m, v = opt_state.get(n, (torch.zeros_like(sd[n]), torch.zeros_like(sd[n])))
        if found_inf:
            new_sd[n], new_opt[n] = sd[n], (m, v)
        else:
            pn, mn, vn = adamw_step(sd[n].float(), gi, m, v, step, lr, betas[0], betas[1], eps, wd)
            new_sd[n], new_opt[n] = pn, (mn, vn)
        new_ema[n] = ema_update(ema_shadow[n], new_sd[n], ema_decay) if ema_shadow is not None else None
    return loss, gnorm, new_sd, new_opt, new_ema


# --------------------------------------------------------------------------------------
# Sampler loops                    (src/testing/ddpm_inference.py, ddpim_inference.py,
#                                   src/training_loops/training_utils.py)
# --------------------------------------------------------------------------------------


def ddim_schedule_t_linear(T: int, steps: int) -> List[int]:
    """ddpim_inference.py:47-53: unique_consecutive(round(linspace(T-1,0,steps))) (+0 if missing)."""
    s = torch.unique_consecutive(torch.linspace(T - 1, 0, steps).round().long())
    out = [int(v) for v in s]
    if out[-1] != 0:
        out.append(0)
    return out


def ddim_schedule_alpha_bar(tb, steps: int) -> List[int]:
    """ddpim_inference.py:55-67: t whose alpha-bar is nearest to 1-u, u in linspace(0,1,steps)."""
    ab = tb["alphas_cumprod"]
    picks = {int((ab - (1.0 - float(u))).abs().argmin()) for u in torch.linspace(0.0, 1.0, steps)}
    out = sorted(picks, reverse=True)
    if out[-1] != 0:
        out.append(0)
    return out


def ddim_sample_indices(T: int, steps: int, schedule: str) -> Tensor:
    """training_utils.py:73-87: steps+1 points, linear / cosine_alpha_bar / karras(p=2) spacing."""
    if schedule == "linear":
        idx = torch.linspace(T - 1, 0, steps + 1)
    elif schedule == "cosine_alpha_bar":
        s = torch.linspace(0, 1, steps + 1)
        idx = (T - 1) * (1 - 0.5 * (1 - torch.cos(torch.pi * s)))
    elif schedule == "karras":
        idx = (T - 1) * (1 - torch.linspace(0, 1, steps + 1) ** 2.0)
    else:
        raise ValueError("schedule inválido")
    return idx.round().clamp(0, T - 1).long()


def ddpm_sample_loop(eps_fn, tb, x: Tensor, noises: Sequence[Tensor], T: Optional[int] = None,
                     clamp_x0=True, dynamic_threshold=None) -> Tensor:
    """ddpm_inference.py:36-38: for i = T-1..0: x = p_sample_step(model, x, full(i)).
    ``noises[k]`` is the k-th randn_like drawn by the loop (identical-noise parity)."""
    T = tb["betas"].shape[0] if T is None else T
    B = x.shape[0]
    for k, i in enumerate(reversed(range(T))):
        t = torch.full((B,), i, dtype=torch.long)
        x = ddpm_step(tb, eps_fn(x, t), x, t, noises[k], clamp_x0, dynamic_threshold)
    return x


def ddim_sample_loop(eps_fn, tb, x: Tensor, schedule: Sequence[int], noises: Sequence[Tensor],
                     eta: float = 0.0, clamp_x0=True, dynamic_threshold=None, clip_x0=True) -> Tensor:
    """ddpim_inference.py:74-87: consecutive pairs of the schedule, clip_x0=True."""
    B = x.shape[0]
    for k in range(len(schedule) - 1):
        t = torch.full((B,), int(schedule[k]), dtype=torch.long)
        tp = torch.full((B,), int(schedule[k + 1]), dtype=torch.long)
        x = ddim_step(tb, eps_fn(x, t), x, t, tp, noises[k], eta, clamp_x0, dynamic_threshold, clip_x0)
    return x


def to_image01(x: Tensor) -> Tensor:
    """ddpm_inference.py:40: (clamp(x,-1,1)+1)/2."""
    return (x.clamp(-1, 1) + 1) * 0.5
