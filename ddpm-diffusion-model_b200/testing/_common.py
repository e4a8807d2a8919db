"""Shared pieces of the two sampler modules (EMA weight swap, grid output, batch sharding)."""
import math
import os
from contextlib import contextmanager

import torch

from .. import dist as _dist
from ..model.difussion_class import to_image01


@contextmanager
def sampling_weights(model, ema):
    """eval() + optional EMA swap, restored afterwards (ddpm_inference.py:22-28,54-56)."""
    was_training = model.training
    model.eval()
    backup = None
    if ema is not None:
        backup = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ema.copy_to(model)
    try:
        yield
    finally:
        if backup is not None:
            model.load_state_dict(backup)
        model.train(was_training)


def initial_noise(n, img_size, device, seed, shard):
    """x_T ~ N(0, I).  With `shard=True` under torch.distributed every rank draws the FULL batch
    from the same seed and keeps its own rows, so the union over ranks is bit-identical to a
    single-GPU run; no communication."""
    if seed is not None:
        torch.manual_seed(seed)
    x = torch.randn(n, 3, img_size, img_size, device=device)
    if shard:
        lo, hi = _dist.shard_range(n)
        x = x[lo:hi].contiguous()
    return x


class StepGraph:
    """One sampler transition x_t -> x_{t-1} (UNet forward + the fused DDPM / DDIM kernel, ~150 launches)
    captured once as a CUDA graph and replayed per step: the reference's loop is launch-bound at small batch
    (two host syncs and ~120 ATen calls per step on top of the model).

    The random draw stays OUTSIDE the graph -- `torch.randn_like` on the default generator, in the same order
    as the reference (difussion_class.py:186,213) -- so seeds reproduce bit-identical noise.
    `step(x, t, t_prev, noise) -> x_next` must be a pure function of its tensor arguments."""

    def __init__(self, step, x: torch.Tensor, with_prev: bool):
        B = x.shape[0]
        dev = x.device
        self.x = x.clone()
        self.t = torch.zeros(B, dtype=torch.long, device=dev)
        self.tp = torch.zeros(B, dtype=torch.long, device=dev) if with_prev else None
        self.noise = torch.zeros_like(x)
        self._step = step
        # eager warm-up on a side stream: builds TMA-free host state (packed weights, schedule handle, pooled
        # activation buffers) so that nothing is allocated or packed inside the capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._call()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._call()

    def _call(self):
        return self._step(self.x, self.t, self.tp, self.noise)

    def run(self, cur: int, prev, draw_noise: bool = True) -> None:
        """Advance the internal state by one transition (state lives in `self.x`)."""
        self.t.fill_(cur)
        if self.tp is not None:
            self.tp.fill_(prev)
        if draw_noise:
            self.noise.copy_(torch.randn_like(self.noise))    # same call (and RNG stream) as difussion_class.py:186,213
        self.graph.replay()
        self.x.copy_(self.out)


def graphs_enabled() -> bool:
    """Opt-in (DDPM_B200_GRAPHS=1).  Measured on B200 (tools/graphdiag.py): a replayed step costs the same as the
    eagerly launched one at every batch size (B=8: 1.60 vs 1.75 ms, B=256: 8.45 vs 8.53 ms) -- the step is bound
    by GPU-side per-kernel latency, not by host launches -- so graphs are not the default."""
    return os.environ.get("DDPM_B200_GRAPHS", "0") == "1"


# ------------------------------------------------------------------------------------------------------------------
# sampler output path: grid assembly + uint8 conversion in one kernel, one pinned D2H copy, PNG
# encoding off the critical path when asked for
# ------------------------------------------------------------------------------------------------------------------
_PNG_LEVEL = int(os.environ.get("DDPM_B200_PNG_LEVEL", "3"))      # zlib level; the decoded pixels are the same at every level
_WRITER = None
_PENDING = []


def _grid_shape(N, H, W, nrow, pad):
    if N == 1:
        return H, W
    xm = min(nrow, N)
    return (H + pad) * math.ceil(N / xm) + pad, (W + pad) * xm + pad


def image_grid(x01: torch.Tensor, nrow: int, pad: int = 2, want_float: bool = True):
    """torchvision.make_grid(x01, nrow, padding=pad) as a device fp32 tensor [3,Hg,Wg] and the uint8 HWC image that
    torchvision.save_image would hand to PIL (pinned host memory, already synchronised).  One kernel + one D2H."""
    from .. import _lib
    if x01.dim() == 3:
        x01 = x01.unsqueeze(0)
    if not (x01.is_cuda and x01.dim() == 4 and x01.shape[1] in (1, 3)):
        raise RuntimeError("ddpm_b200.image_grid expects a CUDA tensor [N,1|3,H,W] (no CPU fallback)")
    x = x01.detach().to(torch.float32).contiguous()
    N, Cc, H, W = x.shape
    Hg, Wg = _grid_shape(N, H, W, nrow, pad)
    grid = torch.empty((3, Hg, Wg), dtype=torch.float32, device=x.device) if want_float else None
    u8 = torch.empty((Hg, Wg, 3), dtype=torch.uint8, device=x.device)
    _lib.call("ddpm_image_grid", x.data_ptr(), N, Cc, H, W, int(nrow), int(pad), grid.data_ptr() if want_float else None,
              u8.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    host = torch.empty((Hg, Wg, 3), dtype=torch.uint8).pin_memory()
    host.copy_(u8, non_blocking=True)
    torch.cuda.current_stream(x.device).synchronize()
    return grid, host


def _encode(arr, path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    if str(path).lower().endswith(".png"):
        from ._png import encode_png
        with open(path, "wb") as f:
            f.write(encode_png(arr, _PNG_LEVEL))
    else:                                                      # other formats: PIL, like torchvision.save_image
        from PIL import Image
        Image.fromarray(arr).save(path)


def write_image(host_u8: torch.Tensor, path: str) -> None:
    """PNG (or whatever the extension says) from an HWC uint8 host tensor.  With DDPM_B200_ASYNC_IO=1 the encoding
    runs on a writer thread so that a sampling sweep is not serialised behind zlib; `flush_image_writes()` (also
    registered with atexit) waits for the files."""
    global _WRITER
    arr = host_u8.numpy()
    if os.environ.get("DDPM_B200_ASYNC_IO", "0") == "1":
        if _WRITER is None:
            import atexit
            from concurrent.futures import ThreadPoolExecutor
            _WRITER = ThreadPoolExecutor(max_workers=2)
            atexit.register(flush_image_writes)
        _PENDING.append(_WRITER.submit(_encode, arr, path))
    else:
        _encode(arr, path)


def flush_image_writes() -> None:
    while _PENDING:
        _PENDING.pop().result()


def save_grid(x01, nrow, out_path, pad=2):
    """vutils.make_grid + vutils.save_image of the reference samplers; returns the float grid like they do."""
    grid, host = image_grid(x01, nrow, pad)
    write_image(host, out_path)
    return grid


def save_each(x01, out_dir):
    """One file per sample (ddpm_inference.py:47-51): one kernel + one D2H for the whole batch (a vertical stack with no
    padding is N contiguous HWC images), then one encode per file."""
    os.makedirs(out_dir, exist_ok=True)
    N, H = x01.shape[0], x01.shape[2]
    _, host = image_grid(x01, 1, 0, want_float=False)
    for i in range(N):
        write_image(host[i * H:(i + 1) * H], os.path.join(out_dir, f"img_{i:03d}.png"))


__all__ = ["sampling_weights", "initial_noise", "save_grid", "save_each", "image_grid", "write_image", "flush_image_writes", "to_image01", "math"]
