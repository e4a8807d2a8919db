"""Drop-in for src/model/unet_backbone.py: same constructors, attribute names, parameter order and
state_dict keys; forward passes run on libddpm_b200 (NHWC, concat-by-layout, fused epilogues)."""
from typing import Optional, Sequence, Set, Tuple

import torch
import torch.nn as nn

from .. import functional as Fn
from .attention import AttnBlock, SinusoidalPosEmb, TimeMLP, group_norm  # noqa: F401


class ResBlock(nn.Module):
    """unet_backbone.py:10-44 -- conv2(drop(silu(gn2(conv1(silu(gn1 x)) + time_proj(t)))) + skip(x)."""

    def __init__(self, in_ch: int, out_ch: int, time_dim: int, dropout: float = 0.0):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.norm1, self.act1 = group_norm(in_ch), nn.SiLU()
        self.conv1 = nn.Conv2d(in_ch, out_ch, kernel_size=3, padding=1)
        self.time_proj = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, out_ch))
        self.norm2, self.act2 = group_norm(out_ch), nn.SiLU()
        self.drop = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.conv2 = nn.Conv2d(out_ch, out_ch, kernel_size=3, padding=1)
        self.skip = nn.Conv2d(in_ch, out_ch, kernel_size=1) if in_ch != out_ch else nn.Identity()

    def forward(self, x: torch.Tensor, t_emb: torch.Tensor) -> torch.Tensor:
        return Fn.resblock_apply(self, x, t_emb)


class Downsample(nn.Module):
    """unet_backbone.py:47-54 -- 3x3 stride-2 conv."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, stride=2, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return Fn.down_apply(self, x)


class Upsample(nn.Module):
    """unet_backbone.py:56-64 -- nearest x2 then 3x3 conv."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return Fn.up_apply(self, x)


class _Level(nn.Module):
    """Plain container (the reference uses a bare nn.Module with .blocks and .down/.up)."""


class UNetDenoiser(nn.Module):
    """unet_backbone.py:68-216.  Registration order (time_mlp, in_conv, out_norm, out_conv, downs,
    mid, ups) is part of the contract: EMA.shadow is positional (ema.py:9-13)."""

    def __init__(self, in_channels: int = 3, base_channels: int = 128,
                 channel_mults: Sequence[int] = (1, 2, 2, 2), num_res_blocks: int = 2,
                 attn_resolutions: Set[int] = frozenset({16, 8}), time_embed_dim: int = 512,
                 dropout: float = 0.0, num_heads: int = 4, head_dim: int = 64, img_resolution: int = 64):
        super().__init__()
        td = time_embed_dim
        self.time_pos_emb = SinusoidalPosEmb(td)
        self.time_mlp = TimeMLP(td, td)
        self.in_conv = nn.Conv2d(in_channels, base_channels, kernel_size=3, padding=1)
        self.out_norm, self.out_act = group_norm(base_channels), nn.SiLU()
        self.out_conv = nn.Conv2d(base_channels, in_channels, kernel_size=3, padding=1)

        def attn(ch):
            return AttnBlock(ch, num_heads=num_heads, head_dim=head_dim)

        # encoder: per level num_res_blocks x (ResBlock [+ AttnBlock]); one skip per level
        self.downs = nn.ModuleList()
        res, ch, skip_ch = img_resolution, base_channels, []
        n_levels = len(channel_mults)
        for li, mult in enumerate(channel_mults):
            blocks = nn.ModuleList()
            for _ in range(num_res_blocks):
                blocks.append(ResBlock(ch, base_channels * mult, td, dropout))
                ch = base_channels * mult
                if res in attn_resolutions:
                    blocks.append(attn(ch))
            skip_ch.append(ch)
            lvl = _Level()
            lvl.blocks = blocks
            lvl.down = Downsample(ch) if li < n_levels - 1 else nn.Identity()
            self.downs.append(lvl)
            if li < n_levels - 1:
                res //= 2

        self.mid = nn.ModuleList([ResBlock(ch, ch, td, dropout),
                                  attn(ch) if res in attn_resolutions else nn.Identity(),
                                  ResBlock(ch, ch, td, dropout)])

        # decoder: concat skip, (num_res_blocks + 1) ResBlocks, no attention; the Upsample conv of
        # level j runs at the channel count coming out of level j-1
        self.ups = nn.ModuleList()
        for j, mult in enumerate(reversed(tuple(channel_mults))):
            out_ch = base_channels * mult
            blocks = nn.ModuleList([ResBlock(ch + skip_ch[n_levels - 1 - j], out_ch, td, dropout)])
            blocks.extend(ResBlock(out_ch, out_ch, td, dropout) for _ in range(num_res_blocks))
            lvl = _Level()
            lvl.blocks = blocks
            lvl.up = Upsample(ch) if j > 0 else nn.Identity()
            self.ups.append(lvl)
            ch = out_ch

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """x (B,C,H,W) any float dtype / memory format, t (B,) int or float -> eps_pred (B,C,H,W)."""
        return Fn.unet_apply(self, x, t)


def build_unet_64x64(in_channels: int = 3, base_channels: int = 128,
                     channel_mults: Tuple[int, ...] = (1, 2, 2, 2), num_res_blocks: int = 2,
                     attn_resolutions: Set[int] = frozenset({16, 8}), time_embed_dim: int = 512,
                     dropout: float = 0.1, num_heads: int = 4, head_dim: int = 64) -> UNetDenoiser:
    """unet_backbone.py:219-240."""
    return UNetDenoiser(in_channels, base_channels, channel_mults, num_res_blocks, attn_resolutions,
                        time_embed_dim, dropout, num_heads, head_dim, img_resolution=64)
