"""Two eval forwards of the low-GPU UNet at B=256 under bf16 autocast (for an ncu launch list of the sampler's hot loop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = build_unet_64x64(**LOW_GPU).to(dev).eval()
diff = Diffusion(T=1000, img_size=64).to(dev)
x = torch.randn(B, 3, 64, 64, device=dev); t = torch.full((B,), 500, device=dev, dtype=torch.long); tp = t - 10
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        x = diff.p_sample_step_ddim(model, x, t, tp, eta=0.0)
torch.cuda.synchronize()
