"""Drop-in for src/training_loops/grad_scaler.py."""
import inspect
from contextlib import contextmanager

import torch

# The reference uses `_DTYPE_MAP` at grad_scaler.py:59 without defining it (NameError on the CUDA
# path; its notebooks define it inline).  The drop-in needs the CUDA path, so the map exists here.
_DTYPE_MAP = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp16": torch.float16,
              "float16": torch.float16, "half": torch.float16}


def make_grad_scaler(device: str = "cuda", enabled: bool = True):
    """grad_scaler.py:5-28: a torch.amp.GradScaler for `device`, or None when AMP is off.  The
    object is the genuine torch one (init scale 65536, growth 2x/2000 steps, backoff 0.5) so
    `scaler.state_dict()` in checkpoints stays compatible; train_one_epoch advances its `_scale` /
    `_growth_tracker` tensors with a fused kernel instead of calling unscale_/step/update."""
    if not enabled:
        return None
    if hasattr(torch, "amp") and hasattr(torch.amp, "GradScaler"):
        try:
            if len(inspect.signature(torch.amp.GradScaler).parameters) >= 1:
                return torch.amp.GradScaler(device if device in ("cuda", "cpu") else "cuda")
            return torch.amp.GradScaler()
        except Exception:
            pass
    if hasattr(torch.cuda, "amp") and hasattr(torch.cuda.amp, "GradScaler"):
        return torch.cuda.amp.GradScaler()
    return None


def _cuda_dtype_supported(dtype: torch.dtype) -> bool:
    return torch.cuda.is_available() and dtype in (torch.bfloat16, torch.float16)


@contextmanager
def autocast_ctx(device: str = "cuda", enabled: bool = True, dtype: str = "bf16", cache_enabled: bool = True):
    """grad_scaler.py:36-78.  Under this context the UNet runs its bf16 kernels (NHWC bf16
    activations, fp32 accumulation and GroupNorm statistics) and returns a bf16 eps_pred."""
    if not enabled:
        yield
        return
    dev = device if isinstance(device, str) else torch.device(device).type
    if dev.startswith("cuda"):
        want = _DTYPE_MAP.get(dtype.lower(), torch.bfloat16)
        use = want if _cuda_dtype_supported(want) else torch.float16
        with torch.amp.autocast(device_type="cuda", dtype=use, cache_enabled=cache_enabled):
            yield
        return
    if dev == "cpu":
        with torch.amp.autocast(device_type="cpu", dtype=torch.bfloat16, cache_enabled=cache_enabled):
            yield
        return
    yield
