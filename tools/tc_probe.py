"""Compare the tcgen05 convolution against the CUDA-core kernel on the GPU for each operand-layout
mode (run under gpurun; prints a table)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_diffusion_model_b200 import _lib, engine

dev = torch.device("cuda", 0)
CASES = [  # N, Cin, Cout, H, W, k
    (1, 16, 16, 8, 8, 3), (2, 32, 32, 8, 8, 3), (2, 64, 96, 16, 16, 3), (2, 96, 96, 64, 64, 3),
    (2, 192, 192, 32, 32, 3), (1, 288, 96, 64, 64, 3), (2, 384, 192, 16, 16, 3), (2, 192, 192, 8, 8, 3),
    (2, 96, 288, 16, 16, 3), (2, 64, 192, 8, 8, 1), (2, 288, 96, 16, 16, 1), (1, 512, 512, 16, 16, 3),
]
modes = [(0, 0, "NS"), (1, 1, "SW32+baseoff"), (1, 0, "SW32")]
if len(sys.argv) > 1:
    modes = [m for m in modes if m[2] in sys.argv[1:]]
for mode, bo, name in modes:
    _lib.lib.ddpm_set_tc_mode(mode, bo)
    for (N, Ci, Co, H, W, k) in CASES:
        torch.manual_seed(1)
        E = engine.Exec(dev, _lib.BF16, False, False)
        w = torch.nn.Parameter(torch.randn(Co, Ci, k, k, device=dev) / (Ci * k * k) ** 0.5)
        b = torch.randn(Co, device=dev)
        wf, _ = E.wcache.get(E, w, _lib.BF16, False)
        x = E.act(N, H, W, Ci); x.interior().normal_()
        r = E.act(N, H, W, Co); r.interior().normal_()
        tb = torch.randn(N, Co, device=dev)
        y_tc = E.act(N, H, W, Co); y_ref = E.act(N, H, W, Co)
        _lib.lib.ddpm_set_force_simt(1)
        engine.conv(E, x, wf, y_ref, k, 1, k // 2, bias=b, tbias=tb, res=r)
        _lib.lib.ddpm_set_force_simt(0)
        engine.conv(E, x, wf, y_tc, k, 1, k // 2, bias=b, tbias=tb, res=r)
        torch.cuda.synchronize()
        a, c = y_tc.buf.t.float(), y_ref.buf.t.float()
        err = float((a - c).norm() / c.norm())
        halo = float(a[:, 0].abs().max() + a[:, -1].abs().max() + a[:, :, 0].abs().max() + a[:, :, -1].abs().max())
        print(f"{name:14s} N{N} {Ci:3d}->{Co:3d} {H}x{W} k{k}: rel err {err:.3e}  halo {halo:.1e}  {'OK' if err < 1e-2 and halo == 0 else 'BAD'}", flush=True)

# ---- weight gradient: tensor-core kernel vs CUDA-core kernel
if "WGRAD" in sys.argv[1:] or len(sys.argv) == 1:
    _lib.lib.ddpm_set_tc_mode(1, 0)
    for (N, Ci, Co, H, W) in [(2, 32, 32, 16, 16), (2, 96, 96, 64, 64), (4, 192, 192, 32, 32), (2, 288, 96, 64, 64),
                              (2, 384, 192, 16, 16), (8, 192, 192, 8, 8), (2, 96, 192, 32, 32), (2, 64, 48, 16, 16),
                              (1, 512, 512, 16, 16), (128, 96, 96, 64, 64)]:
        torch.manual_seed(2)
        E = engine.Exec(dev, _lib.BF16, True, True)
        x = E.act(N, H, W, Ci); x.interior().normal_()
        dy = E.act(N, H, W, Co); dy.interior().normal_()
        w1 = torch.nn.Parameter(torch.zeros(Co, Ci, 3, 3, device=dev)); w2 = torch.nn.Parameter(torch.zeros(Co, Ci, 3, 3, device=dev))
        _lib.lib.ddpm_set_force_simt(1)
        engine.wgrad(E, x, dy, w1, 3, 1, 1)
        _lib.lib.ddpm_set_force_simt(0)
        engine.wgrad(E, x, dy, w2, 3, 1, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            engine.wgrad(E, x, dy, w2, 3, 1, 1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        w2.grad /= 4
        err = float((w2.grad - w1.grad).norm() / w1.grad.norm())
        fl = 2.0 * N * H * W * Co * Ci * 9
        print(f"WGRAD N{N} {Ci:3d}->{Co:3d} {H}x{W}: rel err {err:.3e} {'OK' if err < 1e-2 else 'BAD'}  {fl/dt/1e12:.1f} TFLOP/s ({dt*1e6:.0f} us)", flush=True)
