"""Drop-in for src/model/attention.py (reference lines cited per class).

The leaf modules (`nn.GroupNorm`, `nn.Conv2d`, `nn.Linear`) are kept as *parameter holders* so
`named_parameters()` / `state_dict()` match the reference key for key; their math is executed by
libddpm_b200 through `functional`, never by ATen.
"""
import torch
import torch.nn as nn

from .. import functional as Fn


class SinusoidalPosEmb(nn.Module):
    """attention.py:7-22 -- [sin(t f_i) | cos(t f_i)], f_i = exp(-ln(1e4) i/(half-1)), zero pad if odd."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        return Fn.sinusoid_apply(t, self.dim)


class TimeMLP(nn.Module):
    """attention.py:25-35 -- Linear, SiLU, Linear (keys net.0.*, net.2.*)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, out_dim), nn.SiLU(), nn.Linear(out_dim, out_dim))

    def forward(self, t_emb: torch.Tensor) -> torch.Tensor:
        return Fn.time_mlp_apply(self, t_emb)


def group_norm(channels: int, num_groups: int = 32) -> nn.GroupNorm:
    """attention.py:38-39 -- min(32, C) groups, eps 1e-6, affine."""
    return nn.GroupNorm(min(num_groups, channels), channels, eps=1e-6, affine=True)


class AttnBlock(nn.Module):
    """attention.py:42-74 -- GN -> 1x1 qkv (no bias) -> softmax(q k^T/sqrt(d)) v -> 1x1 proj -> + x.
    inner = heads*head_dim is independent of `channels`."""

    def __init__(self, channels: int, num_heads: int = 4, head_dim: int = 64, p_drop: float = 0.0):
        super().__init__()
        if min(channels, num_heads, head_dim) <= 0:
            raise AssertionError("channels, num_heads and head_dim must be positive")
        self.channels, self.num_heads, self.head_dim = channels, num_heads, head_dim
        self.p_drop = float(p_drop)
        width = num_heads * head_dim
        self.norm = group_norm(channels)
        self.qkv = nn.Conv2d(channels, 3 * width, kernel_size=1, bias=False)
        self.proj = nn.Conv2d(width, channels, kernel_size=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.p_drop > 0.0:
            raise NotImplementedError("attention dropout (p_drop>0) is never enabled by the reference UNet")
        return Fn.attn_apply(self, x)
