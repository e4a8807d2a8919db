"""Where does a ddim_infer_sample call spend its time?  (diagnostic; Kineto timeline of a 20-point schedule at B=256)"""
import os, sys, time, collections, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.testing.ddpim_inference import ddim_infer_sample
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev)
diff = Diffusion(T=1000, img_size=64).to(dev)
out = os.path.join(tempfile.gettempdir(), "ddim_tl.png")
def call():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ddim_infer_sample(model, diff, n=B, img_size=64, device="cuda:0", ema=None, out_path=out, seed=1234, steps=STEPS, eta=0.0)
call(); torch.cuda.synchronize()
t0 = time.perf_counter(); call(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"B={B} steps={STEPS}: {dt*1e3:.1f} ms per call = {dt*1e3/(STEPS-1):.2f} ms per UNet evaluation")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    call(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ker = sorted((e for e in ev if not e.name.startswith("Memset")), key=lambda e: e.time_range.start)
fam = collections.defaultdict(lambda: [0.0, 0])
for e in ker:
    n = e.name.split("<")[0].split("(")[0].replace("void ", "")
    fam[n][0] += e.time_range.end - e.time_range.start; fam[n][1] += 1
iv = sorted((e.time_range.start, e.time_range.end) for e in ker)
cov, cs, ce, gaps = 0.0, iv[0][0], iv[0][1], []
for s, e in iv[1:]:
    if s > ce:
        cov += ce - cs; gaps.append((s - ce, s)); cs, ce = s, e
    else:
        ce = max(ce, e)
cov += ce - cs
span = iv[-1][1] - iv[0][0]
print(f"GPU span {span/1e3:.1f} ms, busy {cov/1e3:.1f} ms, idle {sum(g for g, _ in gaps)/1e3:.1f} ms in {len(gaps)} gaps; {len(ker)} GPU ops = {len(ker)/(STEPS-1):.0f} per evaluation")
for n, (t, c) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {t/1e3:8.3f} ms  x{c:5d}  avg {t/c:7.1f} us  {n}")
big = sorted(gaps, reverse=True)[:6]
print("largest gaps (us):", [round(g, 1) for g, _ in big])
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
cpu.sort(key=lambda e: -(e.time_range.end - e.time_range.start))
for e in cpu[:8]:
    print(f"  host {(e.time_range.end - e.time_range.start)/1e3:8.2f} ms  {e.name[:70]}")
