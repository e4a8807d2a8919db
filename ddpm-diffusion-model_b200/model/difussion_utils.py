"""Drop-in for src/model/difussion_utils.py (schedules + `extract`).

The schedules are built with the same torch ops in the same order as the reference so the ten [T]
tables are bit-identical (they are uploaded once to __constant__ memory by `Diffusion`)."""
import math
from typing import Literal

import torch

ScheduleKind = Literal["linear", "cosine"]


def extract(a: torch.Tensor, t: torch.Tensor, x_shape: torch.Size) -> torch.Tensor:
    """difussion_utils.py:7-14.  Gathers a[t] as (B,1,...,1).  Like the reference, an int64 `t` is
    clamped *in place* to [0, T-1]; float `t` is truncated.  Used by the public
    helper methods only -- the fused kernels do their own clamped lookup in constant memory."""
    idx = t.long().clamp_(0, a.shape[0] - 1)
    return a.gather(0, idx).view((idx.shape[0],) + (1,) * (len(x_shape) - 1))


def beta_schedule_linear(T: int, beta_min: float = 1e-4, beta_max: float = 2e-2) -> torch.Tensor:
    """difussion_utils.py:16-20."""
    return torch.linspace(beta_min, beta_max, T, dtype=torch.float32)


def _alpha_bar_cosine(t: torch.Tensor, s: float = 0.008) -> torch.Tensor:
    """difussion_utils.py:22-29: cos^2(((t+s)/(1+s)) pi/2), floored at 1e-7 before squaring."""
    return torch.cos((math.pi / 2.0) * ((t + s) / (1.0 + s))).clamp(min=1e-7) ** 2


def beta_schedule_cosine(T: int, s: float = 0.008) -> torch.Tensor:
    """difussion_utils.py:32-40."""
    grid = torch.arange(T + 1, dtype=torch.float32) / T
    ab = _alpha_bar_cosine(grid, s=s)
    ab = ab / ab[0]
    return (1 - (ab[1:] / ab[:-1])).clamp(min=1e-8, max=0.999)
