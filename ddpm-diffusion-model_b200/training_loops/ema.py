"""Drop-in for src/training_loops/ema.py.  `shadow` stays a positional list aligned with
`model.parameters()`, but on a CUDA model it is a list of views into one flat buffer and `update`
is a single kernel (the reference launches 2 kernels per tensor: 364 / 650 per update)."""
import torch

from .. import _lib
from ..arena import ensure_arena


def _stream(dev) -> int:
    return _lib.stream_for(dev)


class EMA:
    """ema.py:3-41: shadow_i <- d * shadow_i + (1 - d) * p_i over trainable parameters."""

    def __init__(self, model, decay=0.999, device=None):
        self.decay = decay
        self.device = device
        self.shadow = []
        self._flat = None
        self._arena = None
        arena = ensure_arena(model)
        if arena is not None:
            self._arena = arena
            self._flat = arena.flat.detach().clone()
            self.shadow = arena.views(self._flat)
        else:
            for p in model.parameters():
                self.shadow.append(p.detach().clone() if p.requires_grad else None)

    def _flat_ok(self, model) -> bool:
        ar = getattr(model, "_ddpm_arena", None)
        if self._flat is None or ar is None or ar is not self._arena or not ar.valid():
            return False
        s = self.shadow
        return (len(s) == len(ar.params) and s[0] is not None and s[0].data_ptr() == self._flat.data_ptr()
                and s[-1].data_ptr() == self._flat.data_ptr() + 4 * ar.offsets[-1])

    def _reflatten(self, model) -> bool:
        """Adopt a flat layout for shadows that were loaded / created per tensor."""
        ar = ensure_arena(model)
        if ar is None or len(self.shadow) != len(ar.params) or any(s is None for s in self.shadow):
            return False
        flat = torch.zeros_like(ar.flat)
        views = ar.views(flat)
        with torch.no_grad():
            for v, s in zip(views, self.shadow):
                v.copy_(s.to(device=v.device, dtype=torch.float32))
        self._flat, self._arena, self.shadow = flat, ar, views
        return True

    @torch.no_grad()
    def update(self, model):
        if self._flat_ok(model) or self._reflatten(model):
            _lib.call("ddpm_ema_update", self._flat.data_ptr(), self._arena.flat.data_ptr(), self._arena.numel,
                      float(self.decay), _stream(self._flat.device))
            return
        for i, p in enumerate(model.parameters()):            # generic layout: one kernel per tensor
            if not p.requires_grad:
                continue
            s = self.shadow[i]
            if self.device is not None:
                s = self.shadow[i] = s.to(self.device)
            if not (s.is_cuda and s.dtype == torch.float32 and s.is_contiguous() and p.is_contiguous()
                    and s.data_ptr() % 16 == 0 and p.data_ptr() % 16 == 0):
                raise RuntimeError("ddpm_b200.EMA.update needs contiguous fp32 CUDA parameters (no CPU fallback)")
            _lib.call("ddpm_ema_update", s.data_ptr(), p.data_ptr(), s.numel(), float(self.decay), _stream(s.device))

    @torch.no_grad()
    def copy_to(self, model):
        from ..engine import GLOBAL_WCACHE
        if self._flat_ok(model):
            self._arena.flat.copy_(self._flat)
            GLOBAL_WCACHE.bump()
            return
        for i, p in enumerate(model.parameters()):
            if p.requires_grad:
                p.data.copy_(self.shadow[i].data)
        # `.data.copy_` does not bump Tensor._version: without this the packed bf16/fp32 weight copies of the
        # convolution kernels would stay at the pre-swap weights (the flat branch above does the same)
        GLOBAL_WCACHE.bump()

    @torch.no_grad()
    def state_dict(self):
        return {"decay": self.decay, "shadow": self.shadow}       # live list, like the reference

    @torch.no_grad()
    def load_state_dict(self, state):
        self.decay = state["decay"]
        self.shadow = state["shadow"]
        self._flat = None                                          # re-flattened lazily on update


@torch.no_grad()
def ema_health(ema, model, rel_tol: float = 5.0):
    """ema.py:45-83: (ok, reason, rel_diff) -- resume-time diagnostic, not on the step path."""
    live = [p for p in model.parameters() if p.requires_grad]
    shad = [s for s in getattr(ema, "shadow", []) if s is not None]
    if len(live) != len(shad):
        return (False, "len_mismatch", float("inf"))
    m = torch.cat([p.detach().float().reshape(-1) for p in live])
    e = torch.cat([s.detach().float().reshape(-1).to(m.device) for s in shad])
    if not torch.isfinite(e).all():
        return (False, "nan_or_inf_in_ema", float("inf"))
    mn, en = m.norm().item(), e.norm().item()
    if en < 1e-12:
        return (False, "ema_zero_norm", float("inf"))
    if mn < 1e-12:
        return (False, "model_zero_norm", float("inf"))
    rel = (m - e).norm().item() / (mn + 1e-8)
    return (False, "large_rel_diff", rel) if rel > rel_tol else (True, "ok", rel)


@torch.no_grad()
def ema_reinit_from_model(ema, model):
    """ema.py:87-94."""
    for i, p in enumerate(model.parameters()):
        if p.requires_grad:
            ema.shadow[i].data.copy_(p.data)


def ema_set_decay(ema, new_decay: float):
    """ema.py:96-100."""
    try:
        ema.decay = float(new_decay)
    except Exception:
        pass
