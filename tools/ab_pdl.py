"""A/B of programmatic dependent launch inside one process (ddpm_set_pdl toggled at run time): train step at B=128
and the eval forward at B=256 / B=16, alternating, 3 rounds each.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200 import _lib
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
x = torch.empty(128, 3, 64, 64, device=dev).uniform_(-1, 1); y = torch.zeros(128)

def train(K):
    train_one_epoch(model, diff, [(x, y)] * K, opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)

def timed(fn, K):
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(K); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / K

train(5)
for rnd in range(3):
    for on in (1, 0):
        _lib.lib.ddpm_set_pdl(on)
        print(f"train B=128 pdl={on}: {timed(train, 30):.3f} ms/step", flush=True)
model.eval()
for B in (256, 16):
    xe = torch.randn(B, 3, 64, 64, device=dev); t = torch.full((B,), 500, device=dev, dtype=torch.long)
    def ev(K):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            for _ in range(K):
                model(xe, t)
    ev(5)
    for rnd in range(3):
        for on in (1, 0):
            _lib.lib.ddpm_set_pdl(on)
            print(f"eval B={B} pdl={on}: {timed(ev, 30):.3f} ms/eval", flush=True)
