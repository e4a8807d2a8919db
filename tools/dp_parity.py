"""Data-parallel parity on real GPUs (run under torchrun with N >= 2 ranks, NCCL):
the gradients left in the flat arena after the bucketed, backward-overlapped all-reduce must equal the gradients
of ONE process on the concatenated batch, and one optimiser step must leave every rank
with identical parameters.  fp32 (no autocast), dropout 0, the per-rank t / noise tensors are supplied explicitly.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_parity.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ddpm_diffusion_model_b200 import dist as D
from ddpm_diffusion_model_b200.arena import ensure_arena
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = UNetDenoiser(3, 32, (1, 2, 2), 1, {8}, 64, 0.0, 2, 16, 32).to(dev).train()
d = Diffusion(T=1000).to(dev)
Bper = 4
g = torch.Generator().manual_seed(5)
x_all = torch.empty(world * Bper, 3, 32, 32).uniform_(-1, 1, generator=g).to(dev)
t_all = torch.randint(1, 1000, (world * Bper,), generator=g).to(dev)
n_all = torch.randn(world * Bper, 3, 32, 32, generator=g).to(dev)

arena = ensure_arena(model)
arena.attach_grads(zero=True)
# ---- reference: this process alone on the whole batch
loss_full = d.loss_simple(model, x_all, t_all.clone(), noise=n_all)
loss_full.backward()
g_full = arena.grad.clone()
arena.grad.zero_()
# ---- data parallel: my shard, bucketed all-reduce (4 KB buckets to force many of them) overlapped with backward
sync = D.attach_grad_sync(model, arena, bucket_bytes=64 << 10)
lo, hi = rank * Bper, (rank + 1) * Bper
sync.begin()
loss = d.loss_simple(model, x_all[lo:hi], t_all[lo:hi].clone(), noise=n_all[lo:hi])
loss.backward()
sync.finish()
torch.cuda.synchronize()
rel = float((arena.grad - g_full).norm() / g_full.norm())
mx = float((arena.grad - g_full).abs().max())
# every rank must hold the same averaged gradient
chk = arena.grad.double().sum().reshape(1)
gathered = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(gathered, chk)
same = all(float(a) == float(gathered[0]) for a in gathered)
lt = torch.tensor([float(loss)], device=dev); dist.all_reduce(lt); mean_loss = float(lt) / world
if rank == 0:
    print(f"world={world} buckets={len(sync.buckets)} rel_err(avg grad vs full batch)={rel:.3e} max_abs={mx:.3e} "
          f"identical_across_ranks={same} mean_of_rank_losses={mean_loss:.6f} full_batch_loss={float(loss_full):.6f}")
    assert rel < 1e-5 and same and abs(mean_loss - float(loss_full)) < 1e-5
    print("dp-parity-ok")
dist.destroy_process_group()
