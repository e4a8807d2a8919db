"""Pins oracle/ddpm_oracle.py against fixtures produced by the live reference
(tools/make_golden.py).  CPU only.  Tolerances: bit-exact for the [T] tables and index logic;
1e-6 abs for single elementwise formulas (identical op order); 2e-5 relative L2 for UNet
forward/backward (same fp32 ATen kernels, different composition for attention)."""
import math

import torch
import torchvision.utils as vutils

from oracle import ddpm_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_tables_bit_exact(golden):
    g = golden("tables.pt")
    for key, rec in g.items():
        tb = O.make_tables(**rec["kwargs"])
        assert set(tb) == set(rec["tables"]) == set(O.TABLE_NAMES)
        for k in O.TABLE_NAMES:
            assert torch.equal(tb[k], rec["tables"][k]), (key, k)


def test_reference_own_pins():
    # the only numeric pins in the reference's tests (testing/test_diffusion_utils.py:87-88)
    b = O.betas_linear(1000)
    assert abs(float(b[0]) - 1e-4) < 1e-9 and abs(float(b[-1]) - 2e-2) < 1e-6
    assert bool((b[1:] > b[:-1]).all())
    c = O.betas_cosine(1000)
    assert float(c.min()) >= 1e-8 and float(c.max()) <= 0.999 + 1e-6


def test_elementwise(golden):
    g = golden("elementwise.pt")
    tb = O.make_tables()
    x0, eps, noise, ep, t, tp = (g[k] for k in ("x0", "eps", "noise", "eps_pred", "t", "t_prev"))
    x_t = O.q_sample(tb, x0, t, eps)
    assert torch.equal(x_t, g["q_sample"])
    assert torch.equal(O.q_sample(tb, x0, g["t_float"], eps), g["q_sample_tfloat"])
    assert torch.equal(O.predict_x0(tb, x_t, ep, t), g["predict_x0_clamp"])
    assert torch.equal(O.predict_x0(tb, x_t, ep, t, clamp_x0=False), g["predict_x0_noclamp"])
    assert torch.allclose(O.predict_x0(tb, x_t, ep, t, dynamic_threshold=0.995), g["predict_x0_dyn"], atol=1e-6, rtol=0)
    assert torch.allclose(O.ddpm_step(tb, ep, x_t, t, noise), g["ddpm_step"], atol=1e-6, rtol=0)
    assert torch.allclose(O.ddpm_step(tb, ep, x_t, t, noise, dynamic_threshold=0.995), g["ddpm_step_dyn"], atol=1e-6, rtol=0)
    assert torch.allclose(O.ddpm_step(tb, ep, x_t, t, noise, clamp_x0=False), g["ddpm_step_noclamp"], atol=1e-5, rtol=1e-6)
    assert torch.allclose(O.ddim_step(tb, ep, x_t, t, tp, noise, eta=0.0), g["ddim_step_eta0"], atol=1e-6, rtol=0)
    assert torch.allclose(O.ddim_step(tb, ep, x_t, t, tp, noise, eta=0.5), g["ddim_step_eta05"], atol=1e-6, rtol=0)
    assert torch.allclose(O.ddim_step(tb, ep, x_t, t, tp, noise, eta=1.0, dynamic_threshold=0.995),
                          g["ddim_step_eta1_dyn"], atol=1e-6, rtol=0)
    fn = lambda a, b: ep  # noqa: E731
    assert torch.allclose(O.loss_simple(tb, fn, x0, t, eps), g["loss_simple"], atol=1e-6)
    assert torch.allclose(O.loss_simple(tb, fn, x0, t, eps, g["weight"]), g["loss_simple_weighted"], atol=1e-6)
    assert torch.allclose(O.sinusoidal(t, 64), g["sinusoid_64"], atol=1e-6)
    assert torch.allclose(O.sinusoidal(t, 33), g["sinusoid_33"], atol=1e-6)


def _unet_case(golden, name):
    g = golden(name)
    sd = g["state_dict"]
    if isinstance(sd, str):
        sd = golden(sd)["state_dict"]
    spec = O.UNetSpec(**g["cfg"])
    tb = O.make_tables()
    with torch.no_grad():
        eps = O.unet_forward(sd, spec, g["x_t"], g["t"])
    assert rel(eps, g["eps_pred"]) < 2e-5
    loss, eps2, grads = O.unet_loss_and_grads(sd, spec, tb, g["x0"], g["t"], g["noise"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    assert list(sd.keys()) == g["param_names"]
    # biases that feed a GroupNorm have mathematically zero gradients (1e-8 round-off noise in
    # both implementations): measure every tensor against a floor of 1e-4 x the largest grad norm
    floor = 1e-3 * max(float(v.norm()) for v in g["grads"].values())
    worst = max(float((grads[k].double() - v.double()).norm()) / max(float(v.norm()), floor)
                for k, v in g["grads"].items())
    assert worst < 1e-4, worst


def test_unet_attn(golden):
    _unet_case(golden, "unet_tiny_attn.pt")


def test_unet_noattn_odd_time_dim(golden):
    _unet_case(golden, "unet_tiny_noattn.pt")


def test_unet_other_resolution(golden):
    _unet_case(golden, "unet_tiny_attn_s32.pt")


def test_train_steps(golden):
    g = golden("train_steps.pt")
    spec = O.UNetSpec(**g["cfg"])
    tb = O.make_tables()
    sd = {k: v.clone() for k, v in g["init_state_dict"].items()}
    names = g["param_names"]
    opt, ema = {}, {k: sd[k].clone() for k in names}
    # Adam turns round-off-only gradients (biases in front of a GroupNorm) into +-lr steps whose
    # sign is noise; such "dead" parameters are excluded from the trajectory comparison.
    torch.manual_seed(123)
    xb = g["batches"][0]
    _, _, g0 = O.unet_loss_and_grads(sd, spec, tb, xb, torch.randint(1, 1000, (xb.shape[0],)), torch.randn_like(xb))
    gmax = max(float(v.norm()) for v in g0.values())
    dead = {k for k, v in g0.items() if float(v.norm()) < 1e-5 * gmax}
    assert 0 < len(dead) < len(names) // 3
    torch.manual_seed(g["rng_seed"])
    total = 0.0
    for i, x0 in enumerate(g["batches"]):
        t = torch.randint(1, 1000, (x0.shape[0],))
        noise = torch.randn_like(x0)
        lr = g["base_lr"] * min(1.0, (i + 1) / g["warmup_steps"])
        loss, gnorm, sd, opt, ema = O.train_step(
            sd, spec, tb, x0, t, noise, opt, ema, lr=lr, step=i + 1, betas=g["opt"]["betas"],
            eps=g["opt"]["eps"], wd=g["opt"]["weight_decay"], grad_clip=g["grad_clip"],
            ema_decay=g["ema_decay"])
        total += float(loss)
    assert abs(total / len(g["batches"]) - g["avg_loss"]) < 1e-5
    assert (g["n_batches"], g["n_images"], g["global_step"]) == (3, 12, 3)
    # per tensor: error of the 3-step trajectory relative to the distance travelled (Adam
    # amplifies round-off on individual near-zero-gradient elements, so no elementwise check)
    for k, s in zip(names, g["final_ema"]):
        ref, init = g["final_state_dict"][k], g["init_state_dict"][k]
        if k in dead:
            assert float((sd[k] - ref).abs().max()) <= 3 * 2e-3 * 1.01, k
            continue
        moved = float((ref - init).norm())
        assert float((sd[k] - ref).norm()) <= 2e-3 * moved + 1e-7, k
        assert float((ema[k] - s).norm()) <= 2e-3 * float((s - init).norm()) + 1e-7, k


def _model_fn(sd, spec):
    return lambda x, t: O.unet_forward(sd, spec, x, t)


def test_samplers(golden):
    g = golden("samplers.pt")
    spec = O.UNetSpec(**g["cfg"])
    sd = g["state_dict"]
    tb = O.make_tables(T=g["T"], beta_min=g["beta_min"], beta_max=g["beta_max"])
    fn = _model_fn(sd, spec)
    with torch.no_grad():
        # ddim_infer_sample(steps=6, eta=0, t_linear), n=4
        torch.manual_seed(1234)
        x = torch.randn(4, 3, 16, 16)
        sched = O.ddim_schedule_t_linear(g["T"], 6)
        noises = []

        class Lazy(list):            # draw randn_like in loop order, like the reference does
            def __getitem__(self, i):
                return torch.randn(x.shape)
        out = O.ddim_sample_loop(fn, tb, x, sched, Lazy(), eta=0.0)
        grid = vutils.make_grid(O.to_image01(out), nrow=2, padding=2)
        assert torch.allclose(grid, g["ddim_grid_eta0"], atol=2e-4), float((grid - g["ddim_grid_eta0"]).abs().max())

        torch.manual_seed(4321)
        x = torch.randn(3, 3, 16, 16)
        sched = O.ddim_schedule_alpha_bar(tb, 5)
        out = O.ddim_sample_loop(fn, tb, x, sched, Lazy(), eta=1.0)
        grid = vutils.make_grid(O.to_image01(out), nrow=math.ceil(math.sqrt(3)), padding=2)
        assert torch.allclose(grid, g["ddim_grid_eta1_abar"], atol=2e-4)

        torch.manual_seed(1234)
        x = torch.randn(4, 3, 16, 16)
        out = O.ddpm_sample_loop(fn, tb, x, Lazy())
        grid = vutils.make_grid(O.to_image01(out), nrow=2, padding=2)
        assert torch.allclose(grid, g["ddpm_grid"], atol=2e-4)

        for key, sch, eta in (("ddim_sample_karras", "karras", 0.0), ("ddim_sample_linear", "linear", 0.3)):
            torch.manual_seed(1234)
            x = torch.randn(4, 3, 16, 16)
            ts = [int(v) for v in O.ddim_sample_indices(g["T"], 6, sch)]
            out = O.ddim_sample_loop(fn, tb, x, ts, Lazy(), eta=eta)
            assert torch.allclose(O.to_image01(out), g[key], atol=2e-4), key
