#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the LIVE, UNMODIFIED reference on the CPU.

Run in the authoring container only (needs /root/reference):

    python tools/make_golden.py

The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed as
small fixtures.  They pin ``oracle/ddpm_oracle.py`` (tests/test_oracle_golden.py) and, through
it, the CUDA path.  Nothing here is imported by the product.
"""
import io
import json
import os
import sys
import contextlib

import torch

REF = os.environ.get("DDPM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)

from src.model.difussion_class import Diffusion  # noqa: E402
from src.model.unet_backbone import UNetDenoiser, build_unet_64x64  # noqa: E402
from src.model.attention import SinusoidalPosEmb  # noqa: E402
from src.training_loops.ema import EMA  # noqa: E402
from src.training_loops.train_one_epoch import train_one_epoch  # noqa: E402
from src.testing.ddpim_inference import ddim_infer_sample  # noqa: E402
from src.testing.ddpm_inference import ddpm_infer_sample  # noqa: E402
from src.training_loops.training_utils import ddim_sample  # noqa: E402
import src.training_loops.training_utils as _tu  # noqa: E402

_tu.os = os  # the reference forgets ``import os``; only save paths use it

torch.set_num_threads(4)
os.makedirs(OUT, exist_ok=True)

# GroupNorm(min(32,C)) needs every channel count >= 32 to be a multiple of 32 => base 32 is the floor.
TINY = dict(in_channels=3, base_channels=32, channel_mults=(1, 2), num_res_blocks=1,
            attn_resolutions=[8], time_embed_dim=64, dropout=0.0, num_heads=2, head_dim=16,
            img_resolution=16)
TINY_NOATTN = dict(TINY, attn_resolutions=[], num_res_blocks=2, channel_mults=(1, 1), time_embed_dim=33)
SMALL = dict(TINY, channel_mults=(1, 1), time_embed_dim=32, head_dim=8)


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def build(cfg, seed=0):
    torch.manual_seed(seed)
    kw = dict(cfg)
    kw["attn_resolutions"] = set(kw["attn_resolutions"])
    kw["channel_mults"] = tuple(kw["channel_mults"])
    return UNetDenoiser(**kw)


def tables():
    out = {}
    for name, kw in (("linear", dict(T=1000, schedule="linear")),
                     ("cosine", dict(T=1000, schedule="cosine")),
                     ("linear50", dict(T=50, schedule="linear", beta_min=1e-3, beta_max=5e-2))):
        d = Diffusion(**kw)
        out[name] = {"kwargs": kw, "tables": {k: v.clone() for k, v in d.named_buffers()}}
    save("tables.pt", out)


def elementwise():
    torch.manual_seed(11)
    B, S = 6, 8
    d = Diffusion(T=1000)
    x0 = torch.empty(B, 3, S, S).uniform_(-1, 1)
    eps = torch.randn(B, 3, S, S)
    noise = torch.randn(B, 3, S, S)
    eps_pred = torch.randn(B, 3, S, S) * 1.3
    t = torch.tensor([0, 1, 17, 500, 998, 999])
    tp = torch.tensor([0, 0, 3, 480, 900, 979])
    rec = {"x0": x0, "eps": eps, "noise": noise, "eps_pred": eps_pred, "t": t, "t_prev": tp}
    rec["q_sample"] = d.q_sample(x0, t.clone(), eps)
    x_t = rec["q_sample"]
    rec["predict_x0_clamp"] = d.predict_x0(x_t, eps_pred, t.clone())
    d_nc = Diffusion(T=1000, clamp_x0=False)
    rec["predict_x0_noclamp"] = d_nc.predict_x0(x_t, eps_pred, t.clone())
    d_dt = Diffusion(T=1000, dynamic_threshold=0.995)
    rec["predict_x0_dyn"] = d_dt.predict_x0(x_t, eps_pred, t.clone())
    fn = lambda a, b: eps_pred  # noqa: E731
    rec["ddpm_step"] = d.p_sample_step(fn, x_t, t.clone(), noise=noise)
    rec["ddpm_step_dyn"] = d_dt.p_sample_step(fn, x_t, t.clone(), noise=noise)
    rec["ddpm_step_noclamp"] = d_nc.p_sample_step(fn, x_t, t.clone(), noise=noise)
    rec["ddim_step_eta0"] = d.p_sample_step_ddim(fn, x_t, t.clone(), tp.clone(), eta=0.0, noise=noise)
    rec["ddim_step_eta05"] = d.p_sample_step_ddim(fn, x_t, t.clone(), tp.clone(), eta=0.5, noise=noise)
    rec["ddim_step_eta1_dyn"] = d_dt.p_sample_step_ddim(fn, x_t, t.clone(), tp.clone(), eta=1.0, noise=noise)
    w = torch.linspace(0.5, 1.5, B)
    rec["weight"] = w
    rec["loss_simple"] = d.loss_simple(lambda a, b: eps_pred, x0, t.clone(), noise=eps)
    rec["loss_simple_weighted"] = d.loss_simple(lambda a, b: eps_pred, x0, t.clone(), noise=eps, weight=w)
    # float / out-of-range timesteps go through extract's truncate+clamp (difussion_utils.py:12)
    tf = torch.tensor([-3.0, 0.9, 17.7, 500.2, 2000.0, 999.0])
    rec["t_float"] = tf
    rec["q_sample_tfloat"] = d.q_sample(x0, tf.clone(), eps)
    rec["sinusoid_64"] = SinusoidalPosEmb(64)(t)
    rec["sinusoid_33"] = SinusoidalPosEmb(33)(t)
    save("elementwise.pt", rec)


def unet_case(name, cfg, B, S, seed, store_weights=True):
    model = build(cfg, seed=seed)
    d = Diffusion(T=1000)
    torch.manual_seed(seed + 100)
    x0 = torch.empty(B, 3, S, S).uniform_(-1, 1)
    t = torch.randint(1, 1000, (B,))
    noise = torch.randn(B, 3, S, S)
    model.train()
    loss = d.loss_simple(model, x0, t.clone(), noise=noise)
    loss.backward()
    with torch.no_grad():
        x_t = d.q_sample(x0, t.clone(), noise)
        eps_pred = model(x_t, t)
    save(name, {
        "cfg": cfg, "x0": x0, "t": t, "noise": noise, "x_t": x_t,
        "state_dict": ({k: v.detach().clone() for k, v in model.state_dict().items()}
                       if store_weights else "unet_tiny_attn.pt"),
        "param_names": [k for k, _ in model.named_parameters()],
        "eps_pred": eps_pred, "loss": loss.detach(),
        "grads": {k: p.grad.detach().clone() for i, (k, p) in enumerate(model.named_parameters())
                  if store_weights or i % 7 == 0},
    })


def train_case():
    cfg = SMALL
    model = build(cfg, seed=3)
    init_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    d = Diffusion(T=1000, img_size=16)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, betas=(0.9, 0.999), weight_decay=0.01)
    ema = EMA(model, decay=0.9)
    torch.manual_seed(5)
    batches = [(torch.empty(4, 3, 16, 16).uniform_(-1, 1), torch.zeros(4)) for _ in range(3)]
    torch.manual_seed(77)   # RNG stream consumed by train_one_epoch: randint, randn_like per batch
    with contextlib.redirect_stdout(io.StringIO()):
        avg, nb, ni, gs = train_one_epoch(model, d, batches, opt, ema=ema, device="cpu",
                                          use_autocast=False, grad_clip=0.5, base_lr=2e-3,
                                          warmup_steps=4, global_step=0)
    save("train_steps.pt", {
        "cfg": cfg, "init_state_dict": init_sd, "batches": [b[0] for b in batches], "rng_seed": 77,
        "opt": dict(lr=2e-3, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8),
        "ema_decay": 0.9, "grad_clip": 0.5, "base_lr": 2e-3, "warmup_steps": 4,
        "avg_loss": avg, "n_batches": nb, "n_images": ni, "global_step": gs,
        "final_state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "final_ema": [s.detach().clone() for s in ema.shadow],
        "param_names": [k for k, _ in model.named_parameters()],
    })


def sampler_case():
    cfg = SMALL
    model = build(cfg, seed=9)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    d50 = Diffusion(T=50, beta_min=1e-3, beta_max=5e-2)
    rec = {"cfg": cfg, "state_dict": sd, "T": 50, "beta_min": 1e-3, "beta_max": 5e-2}
    with contextlib.redirect_stdout(io.StringIO()):
        rec["ddim_grid_eta0"] = ddim_infer_sample(model, d50, n=4, img_size=16, device="cpu", seed=1234,
                                                  steps=6, eta=0.0, out_path=os.path.join("/tmp", "g0.png"))
        rec["ddim_grid_eta1_abar"] = ddim_infer_sample(model, d50, n=3, img_size=16, device="cpu", seed=4321,
                                                       steps=5, eta=1.0, schedule_kind="alpha_bar_cosine",
                                                       out_path=os.path.join("/tmp", "g1.png"))
        rec["ddpm_grid"] = ddpm_infer_sample(model, d50, n=4, img_size=16, device="cpu", seed=1234,
                                             out_path=os.path.join("/tmp", "g2.png"))
        rec["ddim_sample_karras"] = ddim_sample(model, d50, n=4, img_size=16, device="cpu", seed=1234,
                                                steps=6, eta=0.0, schedule="karras")
        rec["ddim_sample_linear"] = ddim_sample(model, d50, n=4, img_size=16, device="cpu", seed=1234,
                                                steps=6, eta=0.3, schedule="linear")
    save("samplers.pt", rec)


def param_manifests():
    out = {}
    low = build_unet_64x64(base_channels=96, channel_mults=(1, 2, 2, 2), num_res_blocks=1,
                           attn_resolutions={8}, num_heads=2, head_dim=32, dropout=0.1)
    out["low_gpu"] = [[k, list(p.shape)] for k, p in low.named_parameters()]
    big = UNetDenoiser(3, 128, (1, 1, 2, 2, 4), 2, {16}, 512, 0.1, 4, 64, 256)
    out["celeba256"] = [[k, list(p.shape)] for k, p in big.named_parameters()]
    dflt = build_unet_64x64()
    out["default64"] = [[k, list(p.shape)] for k, p in dflt.named_parameters()]
    with open(os.path.join(OUT, "param_manifests.json"), "w") as f:
        json.dump(out, f)
    print("wrote param_manifests.json",
          {k: (len(v), sum(int(torch.tensor(s).prod()) for _, s in v)) for k, v in out.items()})


if __name__ == "__main__":
    tables()
    elementwise()
    unet_case("unet_tiny_attn.pt", TINY, B=3, S=16, seed=1)
    unet_case("unet_tiny_noattn.pt", TINY_NOATTN, B=2, S=16, seed=2)
    unet_case("unet_tiny_attn_s32.pt", TINY, B=1, S=32, seed=1, store_weights=False)   # built for 16 px, run at 32 px
    train_case()
    sampler_case()
    param_manifests()
