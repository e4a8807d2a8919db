"""Times the UNMODIFIED reference (vendored byte-for-byte into baseline/_ref by tools/vendor_reference.py).

Two uses, both through the reference's own public API and stock code path:
  * host cores  -- `bench.py --impl reference`: `train_one_epoch(device="cpu")` (reference
    src/training_loops/train_one_epoch.py:11; its default `use_autocast=True` = CPU bf16 autocast,
    grad_scaler.py:66-74, and fp32) and `ddim_infer_sample(device="cpu")` (src/testing/ddpim_inference.py:7);
  * the same B200 -- `gpu_eager_baseline` of our bench line: the same calls with device="cuda", bf16 autocast +
    GradScaler + channels_last + cudnn.benchmark, exactly how the reference's notebooks run it (PyTorch eager:
    cuDNN / cuBLAS / SDPA).  This is the kernel-for-kernel bar of SURVEY.md section 2.2.

This module never imports the product package (`ddpm_diffusion_model_b200`) nor `oracle/`: the process that runs
the reference arm maps none of our native code.  The only touch on the reference is defining the module global
`_DTYPE_MAP` that src/training_loops/grad_scaler.py:59 reads but never defines (the reference's notebooks define it
in a cell, full_notebooks/Difussion_Model_CelebHQ.ipynb cell 14) -- needed for device="cuda" only.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")

LOW_GPU = dict(base_channels=96, channel_mults=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions={8},
               num_heads=2, head_dim=32, dropout=0.1)


def available() -> str | None:
    """None if the vendored reference is importable, else a one-line reason."""
    if not os.path.isdir(os.path.join(REF_ROOT, "src", "model")):
        return f"baseline/_ref/src missing (run tools/vendor_reference.py where /root/reference exists)"
    return None


def _import_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in list(sys.modules):
        if name == "src" or name.startswith("src."):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", None) or ""
            if f and not f.startswith(REF_ROOT):
                raise RuntimeError(f"a different `src` package is already imported ({f}); the reference arm needs a clean process")
    import src.training_loops.grad_scaler as gs
    if not hasattr(gs, "_DTYPE_MAP"):
        gs._DTYPE_MAP = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp16": torch.float16, "float16": torch.float16}
    from src.model.difussion_class import Diffusion
    from src.model.unet_backbone import UNetDenoiser, build_unet_64x64
    from src.training_loops.ema import EMA
    from src.training_loops.grad_scaler import make_grad_scaler
    from src.training_loops.train_one_epoch import train_one_epoch
    from src.testing.ddpim_inference import ddim_infer_sample
    from src.testing.ddpm_inference import ddpm_infer_sample
    return dict(Diffusion=Diffusion, UNetDenoiser=UNetDenoiser, build_unet_64x64=build_unet_64x64, EMA=EMA,
                make_grad_scaler=make_grad_scaler, train_one_epoch=train_one_epoch, ddim_infer_sample=ddim_infer_sample,
                ddpm_infer_sample=ddpm_infer_sample)


def manifest_id() -> str:
    try:
        import hashlib
        m = json.load(open(os.path.join(REF_ROOT, "MANIFEST.json")))
        return hashlib.sha256(json.dumps(m["files"], sort_keys=True).encode()).hexdigest()[:16]
    except Exception:
        return "unknown"


def _build(R, config: str, device: str):
    torch.manual_seed(0)
    if config == "celeba256":
        model = R["UNetDenoiser"](3, 128, (1, 1, 2, 2, 4), 2, {16}, 512, 0.1, 4, 64, 256)
        img, wd, decay = 256, 0.005, 0.9997
    else:
        model = R["build_unet_64x64"](**LOW_GPU)
        img, wd, decay = 64, 0.0, 0.9995
    model = model.to(device)
    diff = R["Diffusion"](T=1000, schedule="linear", beta_min=1e-4, beta_max=2e-2, img_size=img).to(device)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, betas=(0.9, 0.999), weight_decay=wd)
    ema = R["EMA"](model, decay=decay)
    return model, diff, opt, ema, img


def set_host_threads() -> int:
    """All host cores, explicitly: torchrun exports OMP_NUM_THREADS=1, which starved the CPU arm at N>1 in round 1."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_train(batch: int, steps: int, warmup: int, use_autocast: bool = True, config: str = "low64"):
    """K optimiser steps of the reference's train_one_epoch on the host.  Returns (img/s, s/step, loss)."""
    R = _import_reference()
    model, diff, opt, ema, img = _build(R, config, "cpu")
    torch.manual_seed(7)
    x0 = torch.empty(batch, 3, img, img).uniform_(-1, 1)
    y = torch.zeros(batch)
    with contextlib.redirect_stdout(io.StringIO()):
        if warmup:
            R["train_one_epoch"](model, diff, [(x0, y)] * warmup, opt, ema=ema, device="cpu", grad_clip=1.0, use_autocast=use_autocast)
        t0 = time.perf_counter()
        avg, nb, ni, _ = R["train_one_epoch"](model, diff, [(x0, y)] * steps, opt, ema=ema, device="cpu", grad_clip=1.0,
                                              use_autocast=use_autocast)
        dt = time.perf_counter() - t0
    assert nb == steps and ni == steps * batch
    return ni / dt, dt / steps, float(avg)


def cpu_ddim(n: int, steps: int = 100, config: str = "low64"):
    """`ddim_infer_sample(steps=..., n=...)` on the host, fp32 (the reference's sampler has no autocast of its own).
    Returns (samples/s, seconds, UNet evaluations)."""
    R = _import_reference()
    model, diff, _, _, img = _build(R, config, "cpu")
    outp = os.path.join(tempfile.gettempdir(), f"ref_ddim_cpu_{os.getpid()}.png")
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        R["ddim_infer_sample"](model, diff, n=n, img_size=img, device="cpu", ema=None, out_path=outp, seed=1234, steps=steps, eta=0.0)
        dt = time.perf_counter() - t0
    return n / dt, dt, steps - 1


def gpu_eager(config: str = "low64", batch: int = 128, steps: int = 10, warmup: int = 3, ddim_batch: int = 256,
              ddim_steps: int = 100, do_ddim: bool = True) -> dict:
    """The unmodified reference on cuda:0 through PyTorch eager, configured like its notebooks
    (full_notebooks/Difussion_Model_Low_GPU.ipynb: bf16 autocast + GradScaler, use_channels_last=True,
    cudnn.benchmark=True).  Device-timed with CUDA events around ONE train_one_epoch call over K resident batches
    (the same unit of work our arm times)."""
    R = _import_reference()
    torch.backends.cudnn.benchmark = True
    dev = "cuda"                                  # grad_scaler.py:58 compares the string with "cuda"
    model, diff, opt, ema, img = _build(R, config, dev)
    scaler = R["make_grad_scaler"](dev, True)
    torch.manual_seed(7)
    x = torch.empty(batch, 3, img, img).uniform_(-1, 1).to(dev)
    y = torch.zeros(batch)
    kw = dict(scaler=scaler, ema=ema, device=dev, grad_clip=1.0, use_autocast=True, use_channels_last=True)
    out = {"impl": "unmodified reference, PyTorch eager (cuDNN/cuBLAS/SDPA), bf16 autocast + GradScaler + channels_last, cudnn.benchmark",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "reference_tree": manifest_id(),
           "config": config, "batch": batch, "img": img}
    with contextlib.redirect_stdout(io.StringIO()):
        R["train_one_epoch"](model, diff, [(x, y)] * max(3, warmup), opt, **kw)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        avg, nb, ni, _ = R["train_one_epoch"](model, diff, [(x, y)] * steps, opt, **kw)
        e.record()
        torch.cuda.synchronize()
    sec = s.elapsed_time(e) / 1e3
    out["train"] = {"value": ni / sec, "unit": "img/s", "ms_per_step": sec / steps * 1e3, "steps": steps, "loss": float(avg),
                    "timed": "one train_one_epoch call over K resident batches, CUDA events"}
    out["peak_mem_mb"] = torch.cuda.max_memory_allocated() / 2**20
    if do_ddim:
        outp = os.path.join(tempfile.gettempdir(), f"ref_ddim_gpu_{os.getpid()}.png")

        def call():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                R["ddim_infer_sample"](model, diff, n=ddim_batch, img_size=img, device=dev, ema=None, out_path=outp, seed=1234,
                                       steps=ddim_steps, eta=0.0)
        with contextlib.redirect_stdout(io.StringIO()):
            model.to(memory_format=torch.channels_last)
            call()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            call()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out["ddim"] = {"value": ddim_batch / dt, "unit": "samples/s", "steps": ddim_steps, "unet_evals": ddim_steps - 1,
                       "batch": ddim_batch, "ms_per_eval": dt / (ddim_steps - 1) * 1e3,
                       "timed": "whole ddim_infer_sample call under bf16 autocast (wall clock, synchronize both sides), incl. grid PNG"}
    return out


def main(argv=None):
    """`python baseline/reference_arm.py gpu [config] [batch] [steps] [ddim_batch] [ddim_steps]` -> one JSON line
    (run by bench.py in a subprocess so that the reference's `src` package and cuDNN autotuning never share a
    process with the product)."""
    argv = sys.argv[1:] if argv is None else argv
    why = available()
    if why:
        print(json.dumps({"unavailable": why}))
        return 0
    mode = argv[0] if argv else "gpu"
    if mode == "gpu":
        config = argv[1] if len(argv) > 1 else "low64"
        batch = int(argv[2]) if len(argv) > 2 else (32 if config == "celeba256" else 128)
        steps = int(argv[3]) if len(argv) > 3 else 10
        ddim_batch = int(argv[4]) if len(argv) > 4 else (16 if config == "celeba256" else 256)
        ddim_steps = int(argv[5]) if len(argv) > 5 else 100
        real = os.dup(1)
        os.dup2(2, 1)
        try:
            res = gpu_eager(config, batch, steps, 3, ddim_batch, ddim_steps, do_ddim=ddim_batch > 0)
        except Exception as ex:                   # noqa: BLE001 - report, do not crash the caller's bench line
            res = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
        sys.stdout.flush()
        os.write(real, (json.dumps(res) + "\n").encode())
        return 0
    raise SystemExit(f"unknown mode {mode}")


if __name__ == "__main__":
    sys.exit(main())
