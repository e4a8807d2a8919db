"""Drop-in for src/training_loops/training_utils.py."""
import os

import torch
from torchvision.utils import save_image as _tv_save_image

from ..model.difussion_class import to_image01
from ..testing._common import image_grid, write_image


def _restore(model, backup):
    if backup is not None:
        model.load_state_dict(backup)


@torch.no_grad()
def sample_ddpm(model, diffusion, n: int, img_size: int = 64, device="cuda", steps: int = None,
                save_path: str = None, return_grid: bool = True, ema=None):
    """training_utils.py:7-29: T (or `steps`) ancestral steps from pure noise.  `ema=` is accepted
    (and ignored, the caller swaps weights) so `train_ddpm(sample_fn=sample_ddpm)` works, which the
    reference's own signature does not allow."""
    model.eval()
    n_steps = diffusion.T if steps is None else steps
    x = torch.randn(n, 3, img_size, img_size, device=device)
    for i in reversed(range(n_steps)):
        t = torch.full((n,), i, device=device, dtype=torch.long)
        x = diffusion.p_sample_step(model, x, t)
    x = to_image01(x)
    grid, host = image_grid(x, int(n ** 0.5), 2)             # make_grid + save_image's uint8 conversion, one kernel
    if save_path is not None:
        write_image(host, save_path)
    return grid if return_grid else x


def save_image_grid(x: torch.Tensor, path: str, nrow: int | None = None):
    """training_utils.py:33-50 (with the `os` import the reference forgot)."""
    if nrow is None:
        nrow = max(1, int(x.size(0) ** 0.5))
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    if x.is_cuda and x.dim() == 4 and x.shape[1] in (1, 3):
        write_image(image_grid(x, nrow, 2, want_float=False)[1], path)
    else:                                                     # host tensors / odd channel counts: torchvision as is
        _tv_save_image(x.detach().float().cpu(), path, nrow=nrow)
    print(f"[OK] Guardado grid en {path}")


def ddim_timesteps(T: int, steps: int, schedule: str, device) -> torch.Tensor:
    """training_utils.py:73-87: steps+1 indices, 'linear' | 'cosine_alpha_bar' | 'karras' (rho=2)."""
    u = torch.linspace(0, 1, steps + 1, device=device)
    if schedule == "linear":
        idx = torch.linspace(T - 1, 0, steps + 1, device=device)
    elif schedule == "cosine_alpha_bar":
        idx = (T - 1) * (1 - 0.5 * (1 - torch.cos(torch.pi * u)))
    elif schedule == "karras":
        idx = (T - 1) * (1 - u ** 2.0)
    else:
        raise ValueError("schedule inválido")
    return idx.round().clamp_(0, T - 1).long()


@torch.no_grad()
def ddim_sample(model, diffusion, *, n=16, img_size=256, device="cuda", ema=None, save_path=None, seed=1234,
                steps=50, eta=0.0, schedule="karras", clip_x0=True):
    """training_utils.py:54-100: `steps` DDIM transitions over steps+1 schedule points."""
    was_training = model.training
    model.eval()
    backup = None
    if ema is not None:
        backup = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ema.copy_to(model)
    if seed is not None:
        torch.manual_seed(seed)
    x = torch.randn(n, 3, img_size, img_size, device=device)
    ts = ddim_timesteps(diffusion.T, steps, schedule, device).tolist()     # one host read, not 2/step
    for i in range(steps):
        t = torch.full((n,), ts[i], device=device, dtype=torch.long)
        tp = torch.full((n,), ts[i + 1], device=device, dtype=torch.long)
        x = diffusion.p_sample_step_ddim(model, x, t, tp, eta=eta, clip_x0=clip_x0, noise=None)
    x = to_image01(x)
    if save_path:
        save_image_grid(x, save_path, nrow=int(n ** 0.5))
    _restore(model, backup)
    model.train(was_training)
    return x


def get_lr(optimizer):
    return optimizer.param_groups[0]["lr"]


def lr_warmup(optimizer, base_lr, step, warmup_steps=1000):
    """training_utils.py:108-114."""
    if warmup_steps is None or warmup_steps <= 0:
        return
    lr = base_lr * min(1.0, (step + 1) / warmup_steps)
    for g in optimizer.param_groups:
        g["lr"] = lr


@torch.no_grad()
def _swap_to_ema_and_sample(model, ema, diffusion, sample_fn, sample_n, img_size, device, out_path):
    """training_utils.py:116-125."""
    backup = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ema.copy_to(model)
    sample_fn(model, diffusion, n=sample_n, img_size=img_size, device=device, save_path=out_path)
    model.load_state_dict(backup)


def compute_grad_norm(model) -> float:
    """training_utils.py:128-133 semantics (||g||_2 over all parameters) with ONE host sync: the
    flat gradient arena is reduced by the param_reduce kernel when available."""
    ar = getattr(model, "_ddpm_arena", None)
    if ar is not None and ar.grads_attached():
        from .. import _lib
        st = torch.zeros(4, dtype=torch.float32, device=ar.grad.device)
        _lib.call("ddpm_param_reduce", ar.grad.data_ptr(), ar.numel, st.data_ptr(),
                  torch.cuda.current_stream(ar.grad.device).cuda_stream)
        return float(st[0].item()) ** 0.5
    sq = [p.grad.detach().float().pow(2).sum() for p in model.parameters() if p.grad is not None]
    return float(torch.stack(sq).sum().item()) ** 0.5 if sq else 0.0


def gpu_mem_mb(device="cuda"):
    if torch.cuda.is_available() and str(device).startswith("cuda"):
        return torch.cuda.memory_allocated() / (1024 ** 2), torch.cuda.memory_reserved() / (1024 ** 2)
    return 0.0, 0.0
