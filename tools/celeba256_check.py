"""CelebA256 UNet (BASELINE configs 3/4) on one B200: tensor-core vs CUDA-core agreement at B=1, then
train-step and eval-forward timings (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_diffusion_model_b200 import _lib
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser
from ddpm_diffusion_model_b200.training_loops.ema import EMA
from ddpm_diffusion_model_b200.training_loops.grad_scaler import make_grad_scaler
from ddpm_diffusion_model_b200.training_loops.train_one_epoch import train_one_epoch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = UNetDenoiser(3, 128, (1, 1, 2, 2, 4), 2, {16}, 512, 0.0, 4, 64, 256).to(dev).train()
print("params", sum(p.numel() for p in model.parameters()))
diff = Diffusion(T=1000, img_size=256).to(dev)
x0 = torch.empty(1, 3, 256, 256, device=dev).uniform_(-1, 1); t = torch.randint(1, 1000, (1,), device=dev); noise = torch.randn_like(x0)
out = []
for force in (1, 0):
    _lib.lib.ddpm_set_force_simt(force)
    for p in model.parameters(): p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = diff.loss_simple(model, x0, t, noise=noise)
    loss.backward(); torch.cuda.synchronize()
    out.append((float(loss), {k: p.grad.clone() for k, p in model.named_parameters()}))
_lib.lib.ddpm_set_force_simt(0)
gmax = max(float(v.norm()) for v in out[0][1].values())
worst = max((float((out[1][1][k] - v).norm()) / max(float(v.norm()), 1e-2 * gmax), k) for k, v in out[0][1].items())
print(f"B=1 bf16 loss simt {out[0][0]:.6f} tc {out[1][0]:.6f}; worst grad rel diff {worst[0]:.3e} at {worst[1]}")
for p in model.parameters(): p.grad = None
model.drop_p = 0.1
opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.005)
ema = EMA(model, decay=0.9997); scaler = make_grad_scaler("cuda", True)
for B in (8, 32):
    x = torch.empty(B, 3, 256, 256, device=dev).uniform_(-1, 1); y = torch.zeros(B)
    for _ in range(2):
        train_one_epoch(model, diff, [(x, y)], opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
    torch.cuda.synchronize(); t0 = time.perf_counter(); K = 4
    for _ in range(K):
        r = train_one_epoch(model, diff, [(x, y)], opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"train 256px B={B}: {dt*1e3:.1f} ms/step {B/dt:.1f} img/s  ({B*1257.3e9/dt/1e12:.0f} TFLOP/s)  loss {r[0]:.4f}  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
model.eval()
for B in (16, 64):
    x = torch.randn(B, 3, 256, 256, device=dev); tt = torch.full((B,), 500, device=dev, dtype=torch.long)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(2): model(x, tt)
        torch.cuda.synchronize(); t0 = time.perf_counter(); K = 5
        for _ in range(K): model(x, tt)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"eval fwd 256px B={B}: {dt*1e3:.1f} ms  ({B*419.1e9/dt/1e12:.0f} TFLOP/s) -> DDPM-1000 {B/(1000*dt):.2f} samples/s", flush=True)
