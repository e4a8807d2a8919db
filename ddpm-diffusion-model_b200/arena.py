"""Flat fp32 arenas for parameters, gradients, Adam moments and the EMA shadow.

The reference keeps 182 / 325 separate tensors and walks them in Python (train_one_epoch.py:94-115,
ema.py:15-23).  Here every `nn.Parameter` becomes a view into ONE contiguous buffer (16-byte aligned
slots), so the optimiser-side pass is two kernel launches regardless of the number of tensors, the
data-parallel all-reduce works on a handful of large buckets, and `state_dict()` / `EMA.shadow` /
`optimizer.state_dict()` still expose per-parameter tensors of the reference's shapes.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

_ALIGN = 4   # elements (16 B) -> float4 kernels never straddle two parameters' slots


class ParamArena:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.offsets: List[int] = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        dev = params[0].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.grad: Optional[torch.Tensor] = None
        self.offset_of: Dict[int, int] = {id(p): o for p, o in zip(params, self.offsets)}

    def valid(self) -> bool:
        """True while every parameter still aliases its slot (a `.to()` / `.float()` breaks it)."""
        base = self.flat.data_ptr()
        for p, o in ((self.params[0], self.offsets[0]), (self.params[-1], self.offsets[-1]),
                     (self.params[len(self.params) // 2], self.offsets[len(self.params) // 2])):
            if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                return False
        return True

    def views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def attach_grads(self, zero: bool) -> torch.Tensor:
        """Make every `p.grad` a view of one flat buffer (kernels accumulate into it)."""
        if self.grad is None:
            self.grad = torch.zeros_like(self.flat)
            zero = False
        elif zero:
            self.grad.zero_()
        if self.grads_attached() and all(p.grad is not None for p in self.params):
            return self.grad                                 # views are still in place: nothing to rebuild (0.5 ms per call)
        gv = self.views(self.grad)
        for p, g in zip(self.params, gv):
            if p.requires_grad:
                p.grad = g
        return self.grad

    def grads_attached(self) -> bool:
        p = self.params[-1]
        return (self.grad is not None and p.grad is not None
                and p.grad.data_ptr() == self.grad.data_ptr() + 4 * self.offsets[-1])


def ensure_arena(model: torch.nn.Module) -> Optional[ParamArena]:
    """Flatten `model`'s parameters (once; re-done if a `.to()` detached them).  Returns None for
    models this pass cannot own (CPU, non-fp32, frozen parameters)."""
    params = list(model.parameters())
    if not params or any((not p.is_cuda) or p.dtype != torch.float32 or not p.requires_grad for p in params):
        return None
    ar = getattr(model, "_ddpm_arena", None)
    if ar is not None and len(ar.params) == len(params) and all(a is b for a, b in zip(ar.params, params)) and ar.valid():
        return ar
    ar = ParamArena(params)
    object.__setattr__(model, "_ddpm_arena", ar)
    return ar
