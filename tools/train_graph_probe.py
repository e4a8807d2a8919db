"""Feasibility / timing probe: ONE whole train step (timestep draw, q_sample, UNet forward + backward incl. the side-stream
weight gradients, the fused optimiser pass, weight repack, grad zeroing) captured as a CUDA graph and replayed, against the
eager loop (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.arena import ensure_arena
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.training_loops.ema import EMA
from ddpm_diffusion_model_b200.training_loops.grad_scaler import autocast_ctx, make_grad_scaler
from ddpm_diffusion_model_b200.training_loops.train_one_epoch import _get_fused, train_one_epoch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev)
diff = Diffusion(T=1000, img_size=64).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
ema = EMA(model, decay=0.9995); scaler = make_grad_scaler("cuda", True)
x = torch.empty(B, 3, 64, 64, device=dev).uniform_(-1, 1); y = torch.zeros(B)
kw = dict(scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
train_one_epoch(model, diff, [(x, y)] * 5, opt, **kw)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); train_one_epoch(model, diff, [(x, y)] * 20, opt, **kw); e.record(); torch.cuda.synchronize()
print(f"eager epoch: {s.elapsed_time(e) / 20:.3f} ms/step", flush=True)

arena = ensure_arena(model)
fused = _get_fused(model, opt, arena)
arena.attach_grads(zero=True)
model.train()
loss_sum = torch.zeros((), device=dev)


def body():
    t = diff.sample_timesteps(B, device=dev)
    with autocast_ctx(device="cuda", enabled=True, dtype="bf16"):
        loss = diff.loss_simple(model, x, t)
    scaler.scale(loss).backward()
    fused.run(scaler, True, 1.0, ema)
    loss_sum.add_(loss.detach().float())


side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    for _ in range(3):
        body()
torch.cuda.current_stream(dev).wait_stream(side)
torch.cuda.synchronize()
t0 = time.perf_counter()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    body()
torch.cuda.synchronize()
print(f"capture: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
loss_sum.zero_()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
l0 = float(loss_sum) / 3
loss_sum.zero_()
t0 = time.perf_counter()
s.record()
for _ in range(20):
    g.replay()
e.record()
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"graph replay: {s.elapsed_time(e) / 20:.3f} ms/step, host {host / 20 * 1e3:.3f} ms/step, mean loss {float(loss_sum) / 20:.4f} (first three {l0:.4f})", flush=True)
