"""Drop-in for src/training_loops/main_train_loop.py: the epoch orchestrator around `train_one_epoch`
(main_train_loop.py:48-231) -- resume with overrides, run header / per-epoch table, EMA-swap sampling, periodic
and final checkpoints.  Host-side control flow only; every FLOP is inside train_one_epoch / the samplers.

Differences, all deliberate:
  * under torch.distributed every rank trains (data parallel, gradients averaged inside train_one_epoch); only
    rank 0 prints, samples and writes checkpoints, and the per-epoch images/s is the whole-job figure;
  * `sample_fn` is called as the reference does (`ema=None` keyword, main_train_loop.py:200-204); the samplers of
    this package accept it, which the reference's own `sample_ddpm` does not;
  * the Colab/Drive copy (`drive_ckpt_dir`) is a plain directory copy -- there is no Colab here; `sys`/`shutil`,
    which the reference forgot to import (main_train_loop.py:17,42), are imported.
"""
import os
import shutil
import time

import torch

from .. import dist as _dist
from .ema import ema_health, ema_reinit_from_model, ema_set_decay
from .grad_scaler import make_grad_scaler
from .train_one_epoch import train_one_epoch


def _fmt_hms(sec: float) -> str:
    m, s = divmod(int(sec), 60)
    h, m = divmod(m, 60)
    return f"{h:d}:{m:02d}:{s:02d}"


def _rule(w=92, ch="─"):
    return ch * w


def _copy_ckpt_fixed(src_path: str, dst_dir: str, fixed_name: str):
    try:
        if not dst_dir:
            return
        os.makedirs(dst_dir, exist_ok=True)
        dst = os.path.join(dst_dir, fixed_name)
        if os.path.exists(dst):
            os.remove(dst)
        shutil.copy2(src_path, dst)
        print(f"└─ [DRIVE]  copiado (fixed) → {dst}")
    except Exception as e:                                   # like the reference: never fail the run on a copy
        print(f"└─ [DRIVE]  ERROR al copiar: {e}")


def train_ddpm(model, diffusion, train_loader, optimizer, ema, device="cuda", epochs=50, base_lr=2e-4,
               warmup_steps=1000, grad_clip=1.0, use_autocast=True, scaler=None, sample_every=5, sample_n=36,
               img_size=64, sample_fn=None, ckpt_dir="checkpoints", run_name="ddpm", save_every=5, save_last=True,
               resume_path=None, ckpt_utils=None, grad_accum_steps: int = 1, use_channels_last: bool = False,
               on_oom: str = "skip", log_every: int = 0, probe_timesteps=None, log_mem: bool = False,
               log_grad_norm: bool = False, sample_seed=1234, sample_steps=None, reset_optimizer_state: bool = False,
               override_lr=None, override_weight_decay=None, override_ema_decay=None, repair_ema_on_resume: bool = False,
               ema_decay_after_repair: float = 0.9995, drive_ckpt_dir=None, copy_fixed_to_drive: bool = True,
               fixed_drive_name: str = "latest_ddpm.pt"):
    """main_train_loop.py:48-231 (same arguments, same defaults, same printed table)."""
    rank, world = _dist.world()
    chief = rank == 0
    say = print if chief else (lambda *a, **k: None)
    if chief:
        os.makedirs(ckpt_dir, exist_ok=True)
    save_ckpt, load_ckpt = ckpt_utils if ckpt_utils is not None else (None, None)
    if scaler is None and use_autocast:
        scaler = make_grad_scaler(device=device, enabled=True)

    # ---- resume (every rank loads the same file so replicas stay identical)
    global_step, start_epoch, resumed = 0, 0, False
    if resume_path and load_ckpt is not None and os.path.exists(resume_path):
        step_loaded, extra = load_ckpt(resume_path, model, optimizer=None if reset_optimizer_state else optimizer,
                                       scaler=scaler, ema=ema, map_location=device)
        if isinstance(extra, dict):
            global_step = int(extra.get("global_step", step_loaded or 0))
            start_epoch = int(extra.get("epoch", 0)) + 1
        say(f"[RESUME] Cargado: {resume_path} | global_step={global_step} | start_epoch={start_epoch}")
        if reset_optimizer_state:
            say("[RESUME] Optimizer: estado NO cargado (reset).")
        if override_lr is not None:
            for g in optimizer.param_groups:
                g["lr"] = float(override_lr)
            say(f"[RESUME] override_lr → {override_lr:.3e}")
        if override_weight_decay is not None:
            for g in optimizer.param_groups:
                g["weight_decay"] = float(override_weight_decay)
            say(f"[RESUME] override_weight_decay → {override_weight_decay:.3e}")
        if override_ema_decay is not None and hasattr(ema, "decay"):
            ema.decay = float(override_ema_decay)
            say(f"[RESUME] override_ema_decay → {override_ema_decay:.6f}")
        resumed = True
        if ema is not None and repair_ema_on_resume:
            ok, reason, rel = ema_health(ema, model, rel_tol=5.0)
            if not ok:
                ema_reinit_from_model(ema, model)
                ema_set_decay(ema, float(ema_decay_after_repair))
                say(f"[RESUME][EMA][AUTO] CKPT EMA inválida ({reason}, rel={rel:.3f}). Reinicializada | decay={ema.decay:.6f}")
            else:
                say(f"[RESUME][EMA][AUTO] CKPT EMA saludable (rel={rel:.3f}). Se conserva.")

    # ---- header
    ema_decay_val = getattr(ema, "decay", None)
    ema_str = f"{ema_decay_val:.6f}" if isinstance(ema_decay_val, (float, int)) else "on"
    say(_rule())
    say(f"DDPM run: {run_name}" + (f"  [data parallel x{world}]" if world > 1 else ""))
    say(f"Device: {device} | autocast: {use_autocast} | EMA: {ema_str} | epochs: {epochs} | base_lr: {base_lr:.2e} | "
        f"warmup_steps: {warmup_steps}")
    if resumed:
        say("Overrides activos al reanudar:", f"reset_opt={reset_optimizer_state}", f"override_lr={override_lr}",
            f"override_wd={override_weight_decay}", f"override_ema={override_ema_decay}", sep=" ")
    say(_rule())
    say(f"{'ep':>3} | {'step':>8} | {'loss':>10} | {'lr':>9} | {'batches':>8} | {'images':>8} | {'imgs/s':>7} | {'time':>8} | {'warmup':>6}")
    say(_rule())

    total_time = 0.0
    for epoch in range(start_epoch, epochs):
        t0 = time.time()
        avg_loss, n_batches, n_images, global_step = train_one_epoch(
            model=model, diffusion=diffusion, dataloader=train_loader, optimizer=optimizer, scaler=scaler, ema=ema,
            device=device, grad_clip=grad_clip, use_autocast=use_autocast, grad_accum_steps=grad_accum_steps,
            use_channels_last=use_channels_last, on_oom=on_oom, base_lr=base_lr, warmup_steps=warmup_steps,
            global_step=global_step, log_every=log_every, probe_timesteps=probe_timesteps, log_mem=log_mem,
            log_grad_norm=log_grad_norm)
        sec = time.time() - t0
        total_time += sec
        ips = (n_images * world / sec) if sec > 0 else 0.0
        lr_now = optimizer.param_groups[0]["lr"]
        warm = 0.0 if not warmup_steps else min(1.0, global_step / float(warmup_steps))
        say(f"{epoch:3d} | {global_step:8d} | {avg_loss:10.5f} | {lr_now:9.2e} | {n_batches:8d} | {n_images * world:8d} | "
            f"{ips:7.1f} | {_fmt_hms(sec):>8} | {int(100 * warm):3d}%")

        last = epoch == epochs - 1
        # ---- samples with a temporary EMA swap (rank 0 only; the other ranks wait at the next collective)
        if chief and sample_fn is not None and (epoch % sample_every == 0 or last):
            out_path = os.path.join(ckpt_dir, f"{run_name}_samples_e{epoch:03d}.png")
            use_ema, rel = False, float("inf")
            if ema is not None:
                ok, _, rel = ema_health(ema, model, rel_tol=2.0)
                use_ema = bool(ok and rel <= 1.0)
            backup = {k: v.detach().clone() for k, v in model.state_dict().items()}
            was_training = model.training
            if use_ema:
                ema.copy_to(model)
            if sample_seed is not None:
                torch.manual_seed(sample_seed)
            # (`sample_steps` is accepted and ignored, exactly like main_train_loop.py:200-204)
            sample_fn(model, diffusion, n=sample_n, img_size=img_size, device=device, save_path=out_path, ema=None)
            model.load_state_dict(backup)
            model.train(was_training)
            from ..engine import GLOBAL_WCACHE
            GLOBAL_WCACHE.bump()
            say(f"└─ [SAMPLE] grid → {out_path} | EMA_used={use_ema} | rel={rel:.3f}")

        # ---- checkpoints
        if chief and save_ckpt is not None and (epoch % save_every == 0 or last):
            path = os.path.join(ckpt_dir, f"{run_name}_e{epoch:03d}.pt")
            save_ckpt(path, model, optimizer, scaler, ema, step=global_step, extra={"epoch": epoch, "global_step": global_step})
            say(f"└─ [CKPT]   saved → {path}")
            if copy_fixed_to_drive and drive_ckpt_dir:
                _copy_ckpt_fixed(path, drive_ckpt_dir, fixed_drive_name)

    if chief and save_last and save_ckpt is not None:
        path = os.path.join(ckpt_dir, f"{run_name}_last.pt")
        save_ckpt(path, model, optimizer, scaler, ema, step=global_step, extra={"epoch": epochs - 1, "global_step": global_step})
        say(f"└─ [CKPT]   saved → {path}")
        if copy_fixed_to_drive and drive_ckpt_dir:
            _copy_ckpt_fixed(path, drive_ckpt_dir, fixed_drive_name)
    say(_rule())
    say(f"Entrenamiento finalizado en {_fmt_hms(total_time)}")
    say(_rule())
