"""Is the step host-bound?  Times the train step and the eval forward at several batch sizes; at tiny
batch the GPU work is negligible, so ms/step ~ host enqueue time (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.training_loops.ema import EMA
from ddpm_diffusion_model_b200.training_loops.grad_scaler import make_grad_scaler
from ddpm_diffusion_model_b200.training_loops.train_one_epoch import train_one_epoch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev)
diff = Diffusion(T=1000, img_size=64).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
ema = EMA(model, decay=0.9995); scaler = make_grad_scaler("cuda", True)
for B in (4, 32, 128):
    x = torch.empty(B, 3, 64, 64, device=dev).uniform_(-1, 1); y = torch.zeros(B)
    for _ in range(3):
        train_one_epoch(model, diff, [(x, y)], opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = 8
    for _ in range(K):
        train_one_epoch(model, diff, [(x, y)], opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"train B={B:4d}: {dt*1e3:7.2f} ms/step  {B/dt:8.0f} img/s", flush=True)
model.eval()
for B in (1, 8, 64, 256):
    x = torch.randn(B, 3, 64, 64, device=dev); t = torch.full((B,), 500, device=dev, dtype=torch.long)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(3):
            model(x, t)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        K = 20
        for _ in range(K):
            model(x, t)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"eval fwd B={B:4d}: {dt*1e3:7.2f} ms/eval  -> DDIM-100 (99 evals) {B/(99*dt):8.1f} samples/s", flush=True)
