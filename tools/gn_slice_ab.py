"""GroupNorm at 256 px: one launch over the whole tensor vs one launch per channel slice (run under gpurun).

At 256 px the streaming kernels are HBM-bound and their second pass misses L2 (32 images x 16.8 MB in flight).  A launch over a
16-64 channel slice keeps the bytes in flight (CTAs resident x share per CTA) inside the 126 MB L2, so phase 2 should hit --
at the price of 32-128 byte runs per pixel in DRAM.  Timing only: the sliced calls use per-slice statistics tensors.

    python tools/gn_slice_ab.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ddpm_diffusion_model_b200 import _lib, engine

dev = torch.device("cuda", 0)
E = engine.Exec(dev, _lib.BF16, True, True, rng=torch.tensor([1, 2], dtype=torch.int64, device=dev))


def timeit(fn, reps=6):
    fn(); fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


class _GN:          # the engine wrappers only read these attributes
    def __init__(self, gn, c0, C, cpg):
        self.num_groups, self.eps = C // cpg, gn.eps
        self.weight, self.bias = gn.weight[c0:c0 + C], gn.bias[c0:c0 + C]


for (N, C, H) in [(32, 128, 256), (32, 128, 128), (32, 256, 128), (32, 256, 64), (64, 128, 256), (16, 128, 256)]:
    torch.manual_seed(0)
    gn = torch.nn.GroupNorm(32, C).to(dev)
    cpg = C // 32
    x = E.act(N, H, H, C); x.interior().normal_()
    o = E.act(N, H, H, C)
    dy = E.act(N, H, H, C); dy.interior().normal_()
    dx = E.act(N, H, H, C)
    nbytes = N * H * H * C * 2
    row = f"{N:3d}x{C:4d}@{H:<3d} ({nbytes / 1e6:6.0f} MB)"
    for sl in (0, 16, 32, 64):
        if sl and (sl >= C or sl % cpg):
            continue
        if sl == 0:
            st = [None]
            def fwd():
                st[0] = engine.gn_fwd(E, x, gn, 1, 0.1, 3, out=o)[1]
            def bwd():
                engine.gn_bwd(E, x, st[0], gn, 1, 0.1, 3, dy, dx, False, dy_scratch=True)
        else:
            parts = [( _GN(gn, c0, sl, cpg), x.slice(c0, sl), o.slice(c0, sl), dy.slice(c0, sl), dx.slice(c0, sl)) for c0 in range(0, C, sl)]
            sts = [None] * len(parts)
            def fwd():
                for i, (g, xs, os_, _, _) in enumerate(parts):
                    sts[i] = engine.gn_fwd(E, xs, g, 1, 0.1, 3 + 64 * i, out=os_)[1]
            def bwd():
                for i, (g, xs, _, dys, dxs) in enumerate(parts):
                    engine.gn_bwd(E, xs, sts[i], g, 1, 0.1, 3 + 64 * i, dys, dxs, False, dy_scratch=True)
        tf = timeit(fwd); tb = timeit(bwd)
        row += f" | slice {sl:3d}: fwd {tf:7.1f} us ({4 * nbytes / tf / 1e6:5.2f} TB/s alg) bwd {tb:7.1f} us ({6 * nbytes / tb / 1e6:5.2f})"
    print(row, flush=True)
