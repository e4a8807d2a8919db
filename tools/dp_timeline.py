"""Where the data-parallel step spends its extra time: Kineto timeline of rank 0 under torchrun (NCCL), one K-step
train_one_epoch.  For the last step it prints every NCCL all-reduce kernel (start / end relative to the step's first
kernel, bytes are the bucket sizes of dist.GradSync), the end of the last backward kernel, the start of
param_reduce_kernel (the first kernel that needs the reduced gradients) and the gap between them = the exposed
all-reduce tail.  Also the per-family busy time next to a single-GPU run of the same build when N = 1.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/dp_timeline.py [B] [K]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.training_loops.ema import EMA
from ddpm_diffusion_model_b200.training_loops.grad_scaler import make_grad_scaler
from ddpm_diffusion_model_b200.training_loops.train_one_epoch import train_one_epoch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev)
diff = Diffusion(T=1000, img_size=64).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
ema = EMA(model, decay=0.9995)
scaler = make_grad_scaler("cuda", True)
torch.manual_seed(7 + rank)
x = torch.empty(B, 3, 64, 64, device=dev).uniform_(-1, 1)
y = torch.zeros(B)


def epoch(k):
    return train_one_epoch(model, diff, [(x, y)] * k, opt, scaler=scaler, ema=ema, device=f"cuda:{local}", grad_clip=1.0)


epoch(4)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); epoch(K); e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / K
if world > 1:
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
if rank == 0:
    gs = getattr(model, "_ddpm_grad_sync", None)
    print(f"world {world}, B={B}/GPU: {ms:.3f} ms/step unprofiled (max over ranks)")
    if gs is not None:
        print("buckets (MB, launch order): " + " ".join(f"{(b - a) * 4 / 2**20:.2f}" for a, b in gs.buckets))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    epoch(K)
    torch.cuda.synchronize()
if rank == 0:
    ev = [v for v in prof.events() if v.device_type == torch.autograd.DeviceType.CUDA and v.time_range.end > v.time_range.start]
    ker = sorted((v for v in ev if not v.name.startswith("Mem")), key=lambda v: v.time_range.start)
    # split into steps at param_update_kernel
    ends = [i for i, v in enumerate(ker) if "param_update_kernel" in v.name]
    lo = ends[-2] + 1 if len(ends) >= 2 else 0
    last = [v for v in ker[lo:ends[-1] + 1]]
    while last and ("pack_batched" in last[0].name or "step_bump" in last[0].name or "scaler_update" in last[0].name or "Fill" in last[0].name):
        last = last[1:]
    t0 = last[0].time_range.start
    fam = collections.defaultdict(lambda: [0.0, 0])
    for v in last:
        n = v.name.split("<")[0].split("(")[0].replace("void ", "")
        fam[n][0] += v.time_range.end - v.time_range.start; fam[n][1] += 1
    span = last[-1].time_range.end - t0
    print(f"last step: span {span / 1e3:.3f} ms, {len(last)} kernels")
    for n, (tt, c) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"  {tt / 1e3:8.3f} ms  x{c:4d}  {n}")
    nccl = [v for v in last if "nccl" in v.name.lower()]
    red = [v for v in last if "param_reduce_kernel" in v.name]
    if red:
        r0 = red[0].time_range.start
        compute = [v for v in last if "nccl" not in v.name.lower() and v.time_range.end <= r0]
        cend = max(v.time_range.end for v in compute)
        print(f"last backward/compute kernel ends at {(cend - t0) / 1e3:.3f} ms; param_reduce starts at {(r0 - t0) / 1e3:.3f} ms "
              f"-> exposed wait {(r0 - cend):.1f} us")
        for v in nccl:
            print(f"   nccl  {(v.time_range.start - t0) / 1e3:8.3f} -> {(v.time_range.end - t0) / 1e3:8.3f} ms  ({v.time_range.end - v.time_range.start:7.1f} us)  {v.name[:60]}")
        busy_n = sum(v.time_range.end - v.time_range.start for v in nccl)
        print(f"   {len(nccl)} NCCL kernels, {busy_n / 1e3:.3f} ms of NCCL kernel time; the last one ends {(r0 - max(v.time_range.end for v in nccl)) if nccl else 0:.1f} us before param_reduce")
if world > 1:
    dist.destroy_process_group()
