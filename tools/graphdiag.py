import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.testing._common import StepGraph
from ddpm_diffusion_model_b200 import _lib, engine
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev).eval()
diff = Diffusion(T=1000, img_size=64).to(dev)
for B in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8,64,256").split(",")]:
    x = torch.randn(B, 3, 64, 64, device=dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        step = lambda x_, t_, tp_, z_: diff.p_sample_step_ddim(model, x_t=x_, t=t_, t_prev=tp_, eta=0.0, clip_x0=True, noise=z_)
        t = torch.full((B,), 500, device=dev); tp = torch.full((B,), 490, device=dev); z = torch.zeros_like(x)
        for _ in range(3): step(x, t, tp, z)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): step(x, t, tp, z)
        torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 20
        m0 = engine.POOL.misses; l0 = _lib.launch_count(reset=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sg = StepGraph(step, x, True)
        torch.cuda.synchronize(); build = time.perf_counter() - t0
        print("  pool misses during build:", engine.POOL.misses - m0, "launches during build:", _lib.launch_count(reset=True))
        for _ in range(3): sg.run(500, 490)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): sg.run(500, 490)
        torch.cuda.synchronize(); rep = (time.perf_counter() - t0) / 20
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); sg.graph.replay(); e.record(); torch.cuda.synchronize()
        print(f"B={B}: eager {eager*1e3:.2f} ms/step, graph build {build*1e3:.1f} ms, replay {rep*1e3:.2f} ms/step, one replay (events) {s.elapsed_time(e):.2f} ms")
