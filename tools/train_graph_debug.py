import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.arena import ensure_arena
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.training_loops.ema import EMA
from ddpm_diffusion_model_b200.training_loops.grad_scaler import autocast_ctx, make_grad_scaler
from ddpm_diffusion_model_b200.training_loops.train_one_epoch import _get_fused, train_one_epoch
mode = sys.argv[1]
os.environ["DDPM_B200_TRAIN_GRAPH"] = "0"
B = 32
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev)
diff = Diffusion(T=1000, img_size=64).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
ema = EMA(model, decay=0.9995); scaler = make_grad_scaler("cuda", True)
x = torch.empty(B, 3, 64, 64, device=dev).uniform_(-1, 1); y = torch.zeros(B)
kw = dict(scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
train_one_epoch(model, diff, [(x, y)] * 4, opt, **kw)
arena = ensure_arena(model); fused = _get_fused(model, opt, arena); arena.attach_grads(zero=True); model.train()
ring = torch.zeros(16, device=dev); idx = torch.zeros(1, dtype=torch.int64, device=dev); sl = torch.zeros((), device=dev)
def body():
    t = diff.sample_timesteps(B, device=dev)
    with autocast_ctx(device="cuda", enabled=True, dtype="bf16"):
        loss = diff.loss_simple(model, x, t)
    scaler.scale(loss).backward()
    fused.run(scaler, True, 1.0, ema)
    return loss.detach().float()
def ringops(l):
    sl.copy_(l); ring.index_copy_(0, idx, sl.reshape(1)); idx.add_(1).remainder_(16)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        if mode == "ring": ringops(torch.ones((), device=dev) * 2)
        elif mode == "body": body()
        elif mode == "both": ringops(body())
        elif mode == "rand": diff.sample_timesteps(B, device=dev)
        elif mode == "fwd":
            t = diff.sample_timesteps(B, device=dev)
            with autocast_ctx(device="cuda", enabled=True, dtype="bf16"):
                loss = diff.loss_simple(model, x, t)
        elif mode == "fwdbwd":
            t = diff.sample_timesteps(B, device=dev)
            with autocast_ctx(device="cuda", enabled=True, dtype="bf16"):
                loss = diff.loss_simple(model, x, t)
            scaler.scale(loss).backward()
    g.replay(); torch.cuda.synchronize()
    print(mode, "OK")
except Exception as ex:
    print(mode, "FAILED:", str(ex).splitlines()[0][:200])
