"""tcgen05 tensor-core kernels against the exact CUDA-core kernels on the same bf16 inputs (both go
through the C ABI; `ddpm_set_force_simt` selects the implementation).  bf16 tolerance: 1e-2 relative
(identical bf16 inputs, fp32 accumulation in both, only the summation order differs => ~1e-5)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def mods():
    from ddpm_diffusion_model_b200 import _lib, engine
    yield _lib, engine
    _lib.lib.ddpm_set_force_simt(0)


def relerr(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


CONV = [  # N, Cin, Cout, H, W, k, stride
    (1, 16, 16, 8, 8, 3, 1), (2, 64, 96, 16, 16, 3, 1), (2, 96, 96, 64, 64, 3, 1), (2, 192, 192, 32, 32, 3, 1),
    (1, 288, 96, 64, 64, 3, 1), (2, 384, 192, 16, 16, 3, 1), (2, 96, 288, 16, 16, 3, 1), (1, 512, 512, 16, 16, 3, 1),
    (2, 64, 192, 8, 8, 1, 1), (2, 288, 96, 16, 16, 1, 1), (2, 96, 96, 32, 32, 3, 2), (2, 192, 192, 16, 16, 3, 2),
    (2, 16, 96, 32, 32, 3, 1), (2, 96, 16, 32, 32, 3, 1),
]


@pytest.mark.parametrize("case", CONV)
def test_conv_tc_matches_simt(mods, case):
    _lib, engine = mods
    N, Ci, Co, H, W, k, s = case
    torch.manual_seed(1)
    E = engine.Exec(dev(), _lib.BF16, False, False)
    w = torch.nn.Parameter(torch.randn(Co, Ci, k, k, device=dev()) / (Ci * k * k) ** 0.5)
    b = torch.randn(Co, device=dev())
    wf, _ = E.wcache.get(E, w, _lib.BF16, False)
    x = E.act(N, H, W, Ci); x.interior().normal_()
    Ho, Wo = H // s, W // s
    r = E.act(N, Ho, Wo, Co); r.interior().normal_()
    tb = torch.randn(N, Co, device=dev())
    y_tc, y_ref = E.act(N, Ho, Wo, Co), E.act(N, Ho, Wo, Co)
    y_tc.interior().normal_(); y_ref.interior().copy_(y_tc.interior())          # exercised by accum
    for accum in (False, True):
        _lib.lib.ddpm_set_force_simt(1)
        engine.conv(E, x, wf, y_ref, k, s, k // 2, bias=b, tbias=tb, res=r, accum=accum)
        _lib.lib.ddpm_set_force_simt(0)
        n0 = _lib.launch_count(reset=True)
        engine.conv(E, x, wf, y_tc, k, s, k // 2, bias=b, tbias=tb, res=r, accum=accum)
        torch.cuda.synchronize()
        assert relerr(y_tc.buf.t, y_ref.buf.t) < 1e-2
        full = y_tc.buf.t.float()
        assert float(full[:, 0].abs().max() + full[:, -1].abs().max() + full[:, :, 0].abs().max() + full[:, :, -1].abs().max()) == 0
        if not accum:
            # independent leg: ATen fp32 on the same bf16-rounded operands (the tcgen05 path must not only agree with the
            # repo's own CUDA-core kernel)
            ref = torch.nn.functional.conv2d(x.interior().float().permute(0, 3, 1, 2), w.detach().bfloat16().float(), b,
                                             stride=s, padding=k // 2)
            ref = ref + tb[:, :, None, None] + r.interior().float().permute(0, 3, 1, 2)
            assert relerr(y_tc.interior().float().permute(0, 3, 1, 2), ref) < 1e-2


FUSED = [  # N, Cmain, Cin2, Cout, H  -- out = conv3x3(a; w) + conv1x1(x2; w2) + bias, the ResBlock's conv2 + skip
    (2, 96, 288, 96, 64), (2, 192, 384, 192, 32), (3, 192, 96, 192, 32), (4, 192, 384, 192, 16), (8, 192, 384, 192, 8),
    (2, 96, 192, 96, 64), (1, 32, 16, 32, 16), (40, 96, 288, 96, 64),
]


@pytest.mark.parametrize("case", FUSED)
def test_conv_with_fused_1x1_second_operand(mods, case):
    """One launch (K = 9*Cmain + Cin2 on the tensor cores) against the two-launch form on the CUDA cores; the second
    operand is a channel slice of a wider buffer, like the concatenated up-path input."""
    _lib, engine = mods
    N, Cm, C2, Co, H = case
    torch.manual_seed(4)
    E = engine.Exec(dev(), _lib.BF16, False, False)
    w = torch.nn.Parameter(torch.randn(Co, Cm, 3, 3, device=dev()) / (Cm * 9) ** 0.5)
    w2 = torch.nn.Parameter(torch.randn(Co, C2, 1, 1, device=dev()) / C2 ** 0.5)
    b = torch.randn(Co, device=dev())
    wf, _ = E.wcache.get(E, w, _lib.BF16, False)
    w2f, _ = E.wcache.get(E, w2, _lib.BF16, False)
    a = E.act(N, H, H, Cm); a.interior().normal_()
    wide = E.act(N, H, H, C2 + 32); wide.interior().normal_()
    x2 = wide.slice(16, C2)
    y_ref, y_tc = E.act(N, H, H, Co), E.act(N, H, H, Co)
    _lib.lib.ddpm_set_force_simt(1)
    engine.conv(E, a, wf, y_ref, 3, 1, 1, bias=b, in2=x2, w2pack=w2f)
    _lib.lib.ddpm_set_force_simt(0)
    n0 = _lib.launch_count(reset=True)
    engine.conv(E, a, wf, y_tc, 3, 1, 1, bias=b, in2=x2, w2pack=w2f)
    torch.cuda.synchronize()
    assert _lib.launch_count(reset=True) == 1                         # really one kernel
    assert relerr(y_tc.buf.t, y_ref.buf.t) < 1e-2
    ref = torch.nn.functional.conv2d(a.interior().float().permute(0, 3, 1, 2), w.detach().bfloat16().float(), b, padding=1) + \
        torch.nn.functional.conv2d(x2.interior().float().permute(0, 3, 1, 2), w2.detach().bfloat16().float())
    assert relerr(y_tc.interior().float().permute(0, 3, 1, 2), ref) < 1e-2
    full = y_tc.buf.t.float()
    assert float(full[:, 0].abs().max() + full[:, -1].abs().max() + full[:, :, 0].abs().max() + full[:, :, -1].abs().max()) == 0


WGRAD = [  # N, Cin, Cout, H, W, k
    (2, 32, 32, 16, 16, 3), (2, 96, 96, 64, 64, 3), (4, 192, 192, 32, 32, 3), (2, 288, 96, 64, 64, 3),
    (2, 384, 192, 16, 16, 3), (8, 192, 192, 8, 8, 3), (2, 96, 192, 32, 32, 3), (1, 512, 512, 16, 16, 3),
    (2, 288, 96, 64, 64, 1), (4, 384, 192, 32, 32, 1), (8, 64, 192, 8, 8, 1), (2, 16, 96, 32, 32, 3), (2, 96, 16, 32, 32, 3),
]


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("case", WGRAD)
def test_wgrad_tc_matches_simt(mods, case, pair):
    """`pair` = 1: the cta_group::2 pair kernel (two output-channel tiles as one M = 256 MMA) where the layer has an even number
    of 128-row tiles; `pair` = 0 forces the single-CTA kernel (experiment bit 10)."""
    _lib, engine = mods
    N, Ci, Co, H, W, k = case
    if pair and ((Co + 127) // 128) % 2:
        pytest.skip("odd number of output-channel tiles: the single-CTA kernel runs either way")
    _lib.lib.ddpm_set_tc_mode(1 | ((0 if pair else 1024) << 4), 0)
    try:
        _wgrad_case(_lib, engine, N, Ci, Co, H, W, k)
    finally:
        _lib.lib.ddpm_set_tc_mode(1, 0)


def _wgrad_case(_lib, engine, N, Ci, Co, H, W, k):
    torch.manual_seed(2)
    E = engine.Exec(dev(), _lib.BF16, True, True)
    x = E.act(N, H, W, Ci); x.interior().normal_()
    dy = E.act(N, H, W, Co); dy.interior().normal_()
    # padded-channel variants: the parameter has fewer channels than the activation buffers
    ci_v = 3 if Ci == 16 else Ci
    co_v = 3 if Co == 16 else Co
    w1 = torch.nn.Parameter(torch.zeros(co_v, ci_v, k, k, device=dev()))
    w2 = torch.nn.Parameter(torch.zeros(co_v, ci_v, k, k, device=dev()))
    b1 = torch.nn.Parameter(torch.zeros(co_v, device=dev()))
    b2 = torch.nn.Parameter(torch.zeros(co_v, device=dev()))
    _lib.lib.ddpm_set_force_simt(1)
    engine.wgrad(E, x, dy, w1, k, 1, k // 2, bias=b1)
    _lib.lib.ddpm_set_force_simt(0)
    engine.wgrad(E, x, dy, w2, k, 1, k // 2, bias=b2)
    engine.wgrad(E, x, dy, w2, k, 1, k // 2, bias=b2)          # accumulates
    torch.cuda.synchronize()
    assert relerr(w2.grad / 2, w1.grad) < 1e-2
    # bias gradient: the ones-operand MMA of the tensor-core kernel vs the column-sum kernel vs torch
    ref_b = dy.interior().float().sum((0, 1, 2))[:co_v]
    assert relerr(b1.grad, ref_b) < 1e-3
    assert relerr(b2.grad / 2, ref_b) < 1e-3
    # independent leg: ATen's fp32 weight gradient of the same bf16-rounded operands
    ref_w = torch.nn.grad.conv2d_weight(x.interior().float().permute(0, 3, 1, 2).contiguous(), (Co, Ci, k, k),
                                        dy.interior().float().permute(0, 3, 1, 2).contiguous(), padding=k // 2)
    assert relerr(w2.grad / 2, ref_w[:co_v, :ci_v]) < 1e-2


def test_downsample_grads_via_zero_upsample(mods):
    _lib, engine = mods
    from ddpm_diffusion_model_b200.model.unet_backbone import Downsample
    torch.manual_seed(3)
    res = []
    for force in (1, 0):
        torch.manual_seed(3)
        mod = Downsample(96).to(dev())
        x = torch.randn(4, 96, 32, 32, device=dev(), requires_grad=True)
        _lib.lib.ddpm_set_force_simt(force)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = mod(x)
        y.backward(torch.ones_like(y) * 0.5 + y.detach() * 0.1)
        res.append((y.detach().float(), x.grad.clone(), mod.conv.weight.grad.clone(), mod.conv.bias.grad.clone()))
    _lib.lib.ddpm_set_force_simt(0)
    for a, b in zip(res[0], res[1]):
        assert relerr(b, a) < 1e-2


def test_low_gpu_step_tc_vs_simt(mods):
    """Whole low-GPU UNet, bf16 autocast, B=4: loss and every parameter gradient, tensor cores vs
    CUDA cores (same bf16 data flow)."""
    _lib, engine = mods
    from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
    from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
    kw = dict(base_channels=96, channel_mults=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions={8}, num_heads=2, head_dim=32, dropout=0.0)
    torch.manual_seed(0)
    model = build_unet_64x64(**kw).to(dev()).train()
    d = Diffusion(T=1000).to(dev())
    x0 = torch.empty(4, 3, 64, 64, device=dev()).uniform_(-1, 1)
    t = torch.randint(1, 1000, (4,), device=dev()); noise = torch.randn_like(x0)
    out = []
    for force in (1, 0):
        _lib.lib.ddpm_set_force_simt(force)
        for p in model.parameters():
            p.grad = None
        n0 = _lib.launch_count(reset=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = d.loss_simple(model, x0, t, noise=noise)
        loss.backward()
        torch.cuda.synchronize()
        out.append((float(loss), {k: p.grad.clone() for k, p in model.named_parameters()}))
    _lib.lib.ddpm_set_force_simt(0)
    assert abs(out[0][0] - out[1][0]) < 2e-3 * abs(out[0][0])
    gmax = max(float(v.norm()) for v in out[0][1].values())
    worst = max(float((out[1][1][k] - v).norm()) / max(float(v.norm()), 1e-2 * gmax) for k, v in out[0][1].items())
    # two bf16 implementations that differ only in fp32 summation order flip bf16 roundings of
    # activations; through ~30 layers that is a few % on the smallest gradient tensors
    assert worst < 6e-2, worst
