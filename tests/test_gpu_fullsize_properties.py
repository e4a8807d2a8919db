"""Size-independent properties checked at BASELINE.json's FULL sizes (low-GPU UNet, B=128, 64x64), where the oracle is
too slow to run: adjointness of the three convolution kernels (fprop / dgrad / wgrad are the same bilinear form),
linearity, GroupNorm's defining invariants and the adjointness of its backward, the q_sample <-> predict_x0 round trip
and the fused optimiser pass against torch's AdamW over the whole 12.68 M-parameter arena.  Everything goes through
the C ABI on the tensor-core path."""
import pytest
import torch

pytestmark = pytest.mark.gpu
B = 128


def dev():
    return torch.device("cuda", 0)


def dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("shape", [(96, 96, 64), (192, 192, 32), (288, 96, 64), (192, 192, 8)])
def test_conv_fprop_dgrad_wgrad_are_adjoint_at_full_batch(shape):
    """<conv(x; W), dY> = <x, dgrad(dY; W)> = <W, wgrad(x, dY)> for the bf16 operands actually used.  The outputs are
    rounded to bf16 (y, dx) or summed in fp32 (dW); rounding errors are zero-mean, so the inner products over
    10^7-10^8 terms agree far tighter than a single element would."""
    from ddpm_diffusion_model_b200 import _lib, engine
    Ci, Co, H = shape
    torch.manual_seed(0)
    E = engine.Exec(dev(), _lib.BF16, True, True)
    w = torch.nn.Parameter((torch.randn(Co, Ci, 3, 3, device=dev()) / (Ci * 9) ** 0.5).bfloat16().float())
    wf, wd = E.wcache.get(E, w, _lib.BF16, True)
    x = E.act(B, H, H, Ci); x.interior().normal_()
    dy = E.act(B, H, H, Co); dy.interior().normal_()
    y = engine.conv(E, x, wf, E.act(B, H, H, Co), 3, 1, 1)
    dx = engine.conv(E, dy, wd, E.act(B, H, H, Ci), 3, 1, 1)
    engine.wgrad(E, x, dy, w, 3, 1, 1)
    torch.cuda.synchronize()
    s_f, s_d, s_w = dot(y.interior(), dy.interior()), dot(x.interior(), dx.interior()), dot(w.detach(), w.grad)
    scale = (float(y.interior().float().norm()) * float(dy.interior().float().norm()))
    assert abs(s_f - s_d) < 2e-4 * scale and abs(s_f - s_w) < 2e-4 * scale, (s_f, s_d, s_w, scale)
    for t in (y, dx):                                       # the zero halo survived
        full = t.buf.t.float()
        assert float(full[:, 0].abs().max() + full[:, -1].abs().max() + full[:, :, 0].abs().max() + full[:, :, -1].abs().max()) == 0


def test_conv_is_linear_in_its_input_at_full_batch():
    from ddpm_diffusion_model_b200 import _lib, engine
    torch.manual_seed(1)
    E = engine.Exec(dev(), _lib.BF16, False, False)
    w = torch.nn.Parameter(torch.randn(96, 96, 3, 3, device=dev()) / 30.0)
    wf, _ = E.wcache.get(E, w, _lib.BF16, False)
    x1 = E.act(B, 64, 64, 96); x1.interior().normal_()
    x2 = E.act(B, 64, 64, 96); x2.interior().normal_()
    xs = E.act(B, 64, 64, 96); xs.interior().copy_(2.0 * x1.interior().float() - 0.5 * x2.interior().float())
    y1 = engine.conv(E, x1, wf, E.act(B, 64, 64, 96), 3, 1, 1).interior().float()
    y2 = engine.conv(E, x2, wf, E.act(B, 64, 64, 96), 3, 1, 1).interior().float()
    # f(2 x1 - x2/2) against 2 f(x1) - f(x2)/2 evaluated on the bf16-rounded combination (powers of two scale exactly)
    xs_exact = (2.0 * x1.interior().float() - 0.5 * x2.interior().float())
    ys = engine.conv(E, xs, wf, E.act(B, 64, 64, 96), 3, 1, 1).interior().float()
    ref = 2.0 * y1 - 0.5 * y2
    # the only differences: bf16 rounding of xs (2^-9 relative per element, zero mean) and of the three outputs
    assert float((xs.interior().float() - xs_exact).abs().max()) <= 2 ** -7 * float(xs_exact.abs().max())
    assert float((ys - ref).norm() / ref.norm()) < 6e-3


def test_groupnorm_invariants_and_backward_adjoint_at_full_batch():
    from ddpm_diffusion_model_b200 import _lib, engine
    torch.manual_seed(2)
    E = engine.Exec(dev(), _lib.BF16, True, True)
    C, H, G = 96, 64, 32
    gn = torch.nn.GroupNorm(G, C).to(dev())
    x = E.act(B, H, H, C); x.interior().normal_().mul_(3.0).add_(0.7)
    y, st = engine.gn_fwd(E, x, gn, 0, 0.0, 0)
    yi = y.interior().float().reshape(B, H * H, G, C // G)
    assert float(yi.mean((1, 3)).abs().max()) < 2e-3                     # every (image, group) has zero mean ...
    assert float((yi.var((1, 3), unbiased=False) - 1).abs().max()) < 6e-3     # ... and unit variance (gamma=1, beta=0)
    # scale invariance: GN(4 x) = GN(x) up to eps (4 is exact in bf16)
    x4 = E.act(B, H, H, C); x4.interior().copy_(x.interior().float() * 4.0)
    y4, _ = engine.gn_fwd(E, x4, gn, 0, 0.0, 0)
    d4 = y4.interior().float() - y.interior().float()
    assert float(d4.norm() / y.interior().float().norm()) < 2e-3 and float((d4.abs() / (1 + y.interior().float().abs())).max()) < 2 ** -6
    # statistics: the stored sums are the fp64 sums of the bf16 inputs
    xi = x.interior().double().reshape(B, H * H, G, C // G)
    assert torch.allclose(st[:, :, 0], xi.sum((1, 3)), rtol=1e-6, atol=1e-3)
    assert torch.allclose(st[:, :, 1], (xi * xi).sum((1, 3)), rtol=1e-6)
    # backward: dx is orthogonal to the two directions GroupNorm removes (per image and group: constants and x itself),
    # and <dy, J v> = <J^T dy, v> for a random direction v (J evaluated by a central difference in fp32 torch)
    gn.weight.data.uniform_(0.5, 1.5); gn.bias.data.uniform_(-0.5, 0.5)
    dy = E.act(B, H, H, C); dy.interior().normal_()
    dyc = dy.interior().float().clone()
    y2, st2 = engine.gn_fwd(E, x, gn, 1, 0.0, 0)
    dx = engine.gn_bwd(E, x, st2, gn, 1, 0.0, 0, dy, E.act(B, H, H, C), False, dy_scratch=True)
    dxi = dx.interior().float().reshape(B, H * H, G, C // G)
    xg = x.interior().float().reshape(B, H * H, G, C // G)
    # exact math: sum_group dx = 0 and sum_group dx * (x - mean) = 0; bf16 rounding of dx leaves ~1e-5 of sum |.|
    assert float((dxi.sum((1, 3)).abs() / dxi.abs().sum((1, 3))).max()) < 1e-3
    xc = xg - xg.mean((1, 3), keepdim=True)
    assert float(((dxi * xc).sum((1, 3)).abs() / (dxi * xc).abs().sum((1, 3))).max()) < 1e-3
    xt = x.interior().float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = torch.nn.functional.silu(torch.nn.functional.group_norm(xt, G, gn.weight, gn.bias, gn.eps))
    ref.backward(dyc.permute(0, 3, 1, 2))
    assert float((dx.interior().float() - xt.grad.permute(0, 2, 3, 1)).norm() / xt.grad.norm()) < 1e-2
    assert float((y2.interior().float() - ref.detach().permute(0, 2, 3, 1)).norm() / ref.norm()) < 6e-3


def test_q_sample_predict_x0_round_trip_at_full_batch():
    from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
    torch.manual_seed(7)
    d = Diffusion(T=1000, img_size=64).to(dev())
    x0 = torch.empty(B, 3, 64, 64, device=dev()).uniform_(-1, 1)
    eps = torch.randn_like(x0)
    t = torch.randint(0, 1000, (B,), device=dev())
    t[0], t[1] = 0, 999
    xt = d.q_sample(x0, t, eps)
    back = d.predict_x0(xt, eps, t)
    # error amplification 1/sqrt(ab_t) <= 160 at t = 999 on fp32 rounding of x_t (|x_t| <= ~5): 160 * 5 * 6e-8
    assert float((back - x0).abs().max()) < 2e-4
    # variance bookkeeping of the forward process: ab_t * E[x0^2] + (1 - ab_t) per sample
    ab = d.alphas_cumprod[t].double()
    want = ab * (x0.double() ** 2).mean((1, 2, 3)) + (1 - ab) * (eps.double() ** 2).mean((1, 2, 3)) \
        + 2 * (ab * (1 - ab)).sqrt() * (x0.double() * eps.double()).mean((1, 2, 3))
    assert torch.allclose((xt.double() ** 2).mean((1, 2, 3)), want, rtol=1e-5)


def test_fused_optimizer_pass_matches_torch_adamw_on_the_whole_arena():
    """unscale + clip + AdamW + EMA over the full 12.68 M-parameter low-GPU UNet against torch.optim.AdamW +
    clip_grad_norm_ + the reference's EMA formula, two steps, same random gradients."""
    import copy
    from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
    from ddpm_diffusion_model_b200.training_loops.ema import EMA
    from ddpm_diffusion_model_b200.training_loops.train_one_epoch import _get_fused
    from ddpm_diffusion_model_b200.arena import ensure_arena
    kw = dict(base_channels=96, channel_mults=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions={8}, num_heads=2, head_dim=32, dropout=0.1)
    torch.manual_seed(0)
    model = build_unet_64x64(**kw).to(dev())
    ref = copy.deepcopy(model)
    assert sum(p.numel() for p in model.parameters()) == 12_680_259
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.01)
    ropt = torch.optim.AdamW(ref.parameters(), lr=2e-4, weight_decay=0.01)
    arena = ensure_arena(model)
    fused = _get_fused(model, opt, arena)
    ema = EMA(model, decay=0.9995)
    shadow = [p.detach().clone() for p in ref.parameters()]
    for step in range(2):
        arena.attach_grads(zero=True)
        g = torch.Generator(device=dev()).manual_seed(10 + step)
        for p, q in zip(model.parameters(), ref.parameters()):
            gr = torch.randn(p.shape, device=dev(), generator=g) * 0.05
            p.grad.copy_(gr); q.grad = gr.clone()
        fused.run(None, False, 1.0, ema)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        ropt.step()
        for s, q in zip(shadow, ref.parameters()):
            s.mul_(0.9995).add_(q.detach(), alpha=1 - 0.9995)
    num = sum(float(((p.detach() - q.detach()).double() ** 2).sum()) for p, q in zip(model.parameters(), ref.parameters())) ** 0.5
    den = sum(float((q.detach().double() ** 2).sum()) for q in ref.parameters()) ** 0.5
    assert num < 1e-6 * den, (num, den)
    e_num = sum(float(((a - b).double() ** 2).sum()) for a, b in zip(ema.shadow, shadow)) ** 0.5
    assert e_num < 1e-6 * den, (e_num, den)
