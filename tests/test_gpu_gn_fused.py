"""GroupNorm (+SiLU) applied inside the convolution's operand path (conv_tc.cu GNA kernels, ddpm_conv_args.gn_ab) --
north_star (1), reference sites attention.py:38-39,61 and unet_backbone.py:37-38,43-44,215.

Three legs per case, all through the C ABI:
  * against ATen fp32 (F.group_norm -> silu -> conv2d on the same bf16 tensors), bf16 tolerance 2e-2 (BASELINE north_star);
  * against the repo's own two-launch form (ddpm_gn_fwd then ddpm_conv): both round the normalised activation to bf16 once
    with the same arithmetic, so the results must agree far inside bf16 resolution (1e-3 relative);
  * the output halo stays zero (the transformed patch's halo rows must not become SiLU(b)).
And the whole UNet: the sampling path (fusion on) against the same model with the fusion off.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def mods():
    from ddpm_diffusion_model_b200 import _lib, engine
    return _lib, engine


def relerr(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


CASES = [  # N, Cin, Cout, H, W, k, groups, act, extra ("", "res", "in2", "tb")
    (2, 96, 96, 64, 64, 3, 32, 1, "tb"),          # conv1 of a 64x64 ResBlock, time bias in the epilogue
    (3, 96, 96, 64, 64, 3, 32, 1, "res"),         # conv2 with the identity skip
    (2, 192, 192, 32, 32, 3, 32, 1, "in2"),       # conv2 + fused 1x1 skip: the second operand must NOT be normalised
    (5, 192, 192, 8, 8, 3, 32, 1, "tb"),          # a patch spans several images (coefficient table with 4 rows)
    (7, 384, 192, 16, 16, 3, 32, 1, ""),          # concat input, odd image count
    (2, 288, 96, 64, 64, 3, 32, 1, "in2"),
    (2, 192, 576, 16, 16, 1, 32, 0, ""),          # attention: GroupNorm without SiLU into the 1x1 qkv projection
    (1, 128, 128, 128, 128, 3, 32, 1, "res"),     # CelebA256-style widths, 4-channel groups
    (40, 96, 96, 64, 64, 3, 32, 1, "tb"),         # enough work for resident weights and several items per CTA pair
    (2, 96, 16, 64, 64, 3, 32, 1, ""),            # head: out_conv with Cout padded to 16
    (4, 512, 768, 16, 16, 1, 32, 0, ""),          # CelebA256 attention qkv
]


@pytest.mark.parametrize("case", CASES)
def test_conv_with_gn_operand_transform(mods, case):
    _lib, engine = mods
    N, Ci, Co, H, W, k, G, act, extra = case
    torch.manual_seed(7)
    E = engine.Exec(dev(), _lib.BF16, False, False)
    gn = torch.nn.GroupNorm(G, Ci).to(dev())
    with torch.no_grad():
        gn.weight.normal_(1.0, 0.3); gn.bias.normal_(0.0, 0.3)
    w = torch.nn.Parameter(torch.randn(Co, Ci, k, k, device=dev()) / (Ci * k * k) ** 0.5)
    b = torch.randn(Co, device=dev())
    wf, _ = E.wcache.get(E, w, _lib.BF16, False)
    wide = E.act(N, H, W, Ci + 16)                               # the operand is a channel slice, like the concat buffers
    x = wide.slice(16, Ci)
    x.interior().normal_(0.3, 1.5)
    kw = {}
    ref_extra = 0.0
    if extra == "tb":
        tb = torch.randn(N, Co, device=dev()); kw["tbias"] = tb; ref_extra = tb[:, :, None, None]
    elif extra == "res":
        r = E.act(N, H, W, Co); r.interior().normal_(); kw["res"] = r; ref_extra = r.interior().float().permute(0, 3, 1, 2)
    elif extra == "in2":
        C2 = 96
        w2 = torch.nn.Parameter(torch.randn(Co, C2, 1, 1, device=dev()) / C2 ** 0.5)
        w2f, _ = E.wcache.get(E, w2, _lib.BF16, False)
        x2 = E.act(N, H, W, C2); x2.interior().normal_()
        kw["in2"], kw["w2pack"] = x2, w2f
        ref_extra = F.conv2d(x2.interior().float().permute(0, 3, 1, 2), w2.detach().bfloat16().float())

    # two-launch form
    a, _ = engine.gn_fwd(E, x, gn, act, 0.0, 0)
    y_two = engine.conv(E, a, wf, E.act(N, H, W, Co), k, 1, k // 2, bias=b, **kw)
    # fused form
    ab = engine.gn_coeffs(E, x, gn)
    _lib.launch_count(reset=True)
    y_f = engine.conv(E, x, wf, E.act(N, H, W, Co), k, 1, k // 2, bias=b, gn_ab=ab, gn_act=act, **kw)
    torch.cuda.synchronize()
    assert _lib.launch_count(reset=True) == 1

    xn = x.interior().float().permute(0, 3, 1, 2)
    z = F.group_norm(xn, G, gn.weight, gn.bias, gn.eps)
    z = F.silu(z) if act else z
    ref = F.conv2d(z.bfloat16().float(), w.detach().bfloat16().float(), b, padding=k // 2) + ref_extra
    got = y_f.interior().float().permute(0, 3, 1, 2)
    assert relerr(got, ref) < 2e-2
    assert relerr(y_f.buf.t, y_two.buf.t) < 1e-3
    full = y_f.buf.t.float()
    assert float(full[:, 0].abs().max() + full[:, -1].abs().max() + full[:, :, 0].abs().max() + full[:, :, -1].abs().max()) == 0
    # the input (and its neighbours in the wide buffer) is untouched
    assert float(wide.buf.t[..., :16].float().abs().max()) == 0


def test_gn_coeffs_match_group_norm(mods):
    _lib, engine = mods
    torch.manual_seed(3)
    E = engine.Exec(dev(), _lib.BF16, False, False)
    for (N, C, H, G) in [(3, 96, 64, 32), (5, 192, 8, 32), (2, 512, 16, 32), (2, 288, 32, 32)]:
        gn = torch.nn.GroupNorm(G, C).to(dev())
        with torch.no_grad():
            gn.weight.normal_(1.0, 0.3); gn.bias.normal_(0.0, 0.3)
        x = E.act(N, H, H, C); x.interior().normal_(0.5, 2.0)
        ab = engine.gn_coeffs(E, x, gn)
        xn = x.interior().float().permute(0, 3, 1, 2)
        ref = F.group_norm(xn, G, gn.weight, gn.bias, gn.eps)
        got = xn * ab[:, 0, :, None, None] + ab[:, 1, :, None, None]
        assert float((got - ref).abs().max()) < 1e-4 * float(ref.abs().max())


def test_gn_transform_rejected_off_the_tensor_core_path(mods):
    """No silent unfused result: the CUDA-core kernels do not implement gn_ab, so the ABI refuses."""
    _lib, engine = mods
    E = engine.Exec(dev(), _lib.F32, False, False)
    gn = torch.nn.GroupNorm(4, 16).to(dev())
    w = torch.nn.Parameter(torch.randn(16, 16, 3, 3, device=dev()))
    wf, _ = E.wcache.get(E, w, _lib.F32, False)
    x = E.act(1, 8, 8, 16); x.interior().normal_()
    ab = engine.gn_coeffs(E, x, gn)
    with pytest.raises(RuntimeError):
        engine.conv(E, x, wf, E.act(1, 8, 8, 16), 3, 1, 1, gn_ab=ab, gn_act=1)


@pytest.mark.parametrize("which", ["low64", "celeba256"])
def test_unet_eval_with_fused_groupnorm(mods, which):
    """Sampling-path forward (eval, bf16 autocast, no grad): fusion on vs off on the same weights and inputs, and both vs the
    fp32 forward of the same model (bf16 contract 2e-2)."""
    _lib, engine = mods
    from bench import LOW_GPU
    from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser, build_unet_64x64
    torch.manual_seed(11)
    if which == "low64":
        model, B, S = build_unet_64x64(**LOW_GPU).to(dev()).eval(), 3, 64
    else:
        model = UNetDenoiser(3, 128, (1, 1, 2, 2, 4), 2, {16}, 512, 0.0, 4, 64, 256).to(dev()).eval()
        B, S = 1, 256
    x = torch.randn(B, 3, S, S, device=dev())
    t = torch.randint(1, 1000, (B,), device=dev())
    outs = {}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        model(x, t)                                                       # packs the weights (launches that are not part of a forward)
    for flag in ("0", "1"):
        os.environ["DDPM_B200_FUSE_GN"] = flag
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                _lib.launch_count(reset=True)
                outs[flag] = model(x, t).float()
                torch.cuda.synchronize()
                outs["n" + flag] = _lib.launch_count(reset=True)
        finally:
            os.environ.pop("DDPM_B200_FUSE_GN", None)
    with torch.no_grad():
        ref = model(x, t).float()
    assert outs["n1"] == outs["n0"]                                     # a statistics launch replaces each GroupNorm launch
    assert relerr(outs["1"], outs["0"]) < 5e-3
    assert relerr(outs["1"], ref) < 2e-2
    assert relerr(outs["0"], ref) < 2e-2
