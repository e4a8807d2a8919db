"""Drop-in for src/training_loops/chekpoints.py: the same on-disk dictionary
`{"model", "optimizer", "scaler", "ema", "step", ["extra"]}` (chekpoints.py:4-13), so a file written here loads in
the reference and a reference checkpoint loads here.

What is different underneath: parameters, Adam moments and the EMA shadow live in flat arenas.  Saving detaches
and clones every tensor (a view of a flat arena would otherwise drag the whole 50-250 MB storage along once per
tensor into `torch.save`); loading copies the values back INTO the arenas, so the fused optimiser pass keeps
owning the memory.  Unlike the reference, `scaler=None` / `ema=None` are accepted (chekpoints.py:8 crashes on a
missing scaler); the corresponding keys are then simply absent."""
import torch


def _own(obj):
    """Deep copy with every tensor detached and cloned to its own storage."""
    if torch.is_tensor(obj):
        return obj.detach().clone()
    if isinstance(obj, dict):
        return {k: _own(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_own(v) for v in obj)
    return obj


def save_ckpt(path, model, optimizer, scaler, ema, step: int, extra: dict = None):
    """chekpoints.py:4-13."""
    state = {"model": _own(model.state_dict()), "optimizer": _own(optimizer.state_dict()), "step": step}
    if scaler is not None:
        state["scaler"] = scaler.state_dict()
    if ema is not None:
        sd = ema.state_dict()
        state["ema"] = {"decay": sd["decay"], "shadow": _own(sd["shadow"])}
    if extra:
        state["extra"] = extra
    torch.save(state, path)


def load_ckpt(path, model, optimizer=None, scaler=None, ema=None, map_location="cuda"):
    """chekpoints.py:16-25 -> (step, extra)."""
    state = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(state["model"])                 # copies into the (possibly flattened) parameters
    if optimizer is not None and "optimizer" in state:
        optimizer.load_state_dict(state["optimizer"])     # new per-tensor state; re-adopted by the fused step
        fs = getattr(model, "_ddpm_fused_step", None)
        if fs is not None and fs.opt is optimizer:
            object.__setattr__(model, "_ddpm_fused_step", None)
    if scaler is not None and "scaler" in state:
        scaler.load_state_dict(state["scaler"])
    if ema is not None and "ema" in state:
        ema.load_state_dict(state["ema"])
    from ..engine import GLOBAL_WCACHE
    GLOBAL_WCACHE.bump()                                  # packed weight copies are stale
    return state.get("step", 0), state.get("extra", {})
