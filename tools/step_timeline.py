"""In-situ timeline of the train step (diagnostic, not a bench value): Kineto/CUPTI kernel records of K steady-state
steps -> per-kernel-family busy time, GPU idle gaps between consecutive kernels on the main stream, and host enqueue
time.  Run under gpurun:  python tools/step_timeline.py [B] > gpurun_out/timeline.txt"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.training_loops.ema import EMA
from ddpm_diffusion_model_b200.training_loops.grad_scaler import make_grad_scaler
from ddpm_diffusion_model_b200.training_loops.train_one_epoch import train_one_epoch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
CFG = sys.argv[3] if len(sys.argv) > 3 else "low64"          # or celeba256 (B = 32)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if CFG == "celeba256":
    from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser
    model = UNetDenoiser(3, 128, (1, 1, 2, 2, 4), 2, {16}, 512, 0.1, 4, 64, 256).to(dev)
    IMG = 256
else:
    model = build_unet_64x64(**LOW_GPU).to(dev)
    IMG = 64
diff = Diffusion(T=1000, img_size=IMG).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
ema = EMA(model, decay=0.9995); scaler = make_grad_scaler("cuda", True)
x = torch.empty(B, 3, IMG, IMG, device=dev).uniform_(-1, 1); y = torch.zeros(B)
step = lambda: train_one_epoch(model, diff, [(x, y)], opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
epoch = lambda: train_one_epoch(model, diff, [(x, y)] * K, opt, scaler=scaler, ema=ema, device="cuda:0", grad_clip=1.0)
for _ in range(4):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
epoch()
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / K
print(f"B={B}: {t_all*1e3:.2f} ms/step unprofiled (one train_one_epoch call over {K} batches)")

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    epoch()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ker = [e for e in ev if not e.name.startswith("Memcpy") and not e.name.startswith("Memset")]
ker.sort(key=lambda e: e.time_range.start)
span = (ker[-1].time_range.end - ker[0].time_range.start) / K
fam = collections.defaultdict(lambda: [0.0, 0])
for e in ker:
    n = e.name.split("<")[0].split("(")[0].replace("void ", "")
    fam[n][0] += e.time_range.end - e.time_range.start; fam[n][1] += 1
busy = sum(v[0] for v in fam.values()) / K
print(f"profiled: span {span/1e3:.2f} ms/step, sum of kernel durations {busy/1e3:.2f} ms/step, {len(ker)//K} kernels/step")
for n, (t, c) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {t/K/1e3:8.3f} ms  x{c//K:4d}  avg {t/c:7.1f} us  {n}")
# idle gaps: union of kernel intervals over all streams vs span
iv = sorted((e.time_range.start, e.time_range.end) for e in ker)
covered, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
for s, e in iv[1:]:
    if s > cur_e:
        covered += cur_e - cur_s; gaps.append(s - cur_e); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
covered += cur_e - cur_s
print(f"GPU busy (union over streams) {covered/K/1e3:.2f} ms/step; idle {sum(gaps)/K/1e3:.2f} ms/step in {len(gaps)//K} gaps/step "
      f"(median {sorted(gaps)[len(gaps)//2]:.1f} us, >10us: {sum(1 for g in gaps if g > 10)//K}/step totalling {sum(g for g in gaps if g > 10)/K/1e3:.2f} ms)")
# the largest gaps and what ran just before/after
edges = []
cur_e, last = iv[0][1], ker[0]
for e in ker[1:]:
    if e.time_range.start > cur_e:
        edges.append((e.time_range.start - cur_e, last.name[:50], e.name[:50]))
    if e.time_range.end >= cur_e:
        cur_e, last = e.time_range.end, e
for g, a, b in sorted(edges, reverse=True)[:12]:
    print(f"   gap {g:7.1f} us  after {a}  before {b}")
# per-launch durations of the tensor-core kernels of the LAST step, in launch order (diagnostic: which layers cost what in situ)
if os.environ.get("TIMELINE_SEQ"):
    per = len(ker) // K
    last = ker[-per:]
    for fam_name in ("conv_tc2_kernel", "wgrad_tc_kernel", "gn_bwd_kernel", "gn_fwd_kernel"):
        seq = [round(e.time_range.end - e.time_range.start, 1) for e in last if fam_name in e.name]
        print(fam_name, len(seq), seq)
