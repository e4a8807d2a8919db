"""BASELINE.json configs[3] and [4] on one GPU: DDIM 50/100-step sampling sweep at 64 and 256 px over batch sizes, and the
CelebA256 DDPM-1000 ancestral sampler at B=64, all through the public sampler API under bf16 autocast (run under gpurun).
Prints one line per case and writes gpurun_out/config_sweep.json."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser, build_unet_64x64
from ddpm_diffusion_model_b200.testing.ddpim_inference import ddim_infer_sample
from ddpm_diffusion_model_b200.testing.ddpm_inference import ddpm_infer_sample
dev = torch.device("cuda", 0)
rows = []
quick = "--quick" in sys.argv


def timed(fn):
    with torch.autocast("cuda", dtype=torch.bfloat16), contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return time.perf_counter() - t0


for px in (64, 256):
    torch.manual_seed(0)
    if px == 64:
        model = build_unet_64x64(**LOW_GPU).to(dev).eval()
        batches = (1, 8, 64, 512)
    else:
        model = UNetDenoiser(3, 128, (1, 1, 2, 2, 4), 2, {16}, 512, 0.1, 4, 64, 256).to(dev).eval()
        batches = (1, 8, 64) if quick else (1, 8, 64, 128)
    diff = Diffusion(T=1000, img_size=px).to(dev)
    for steps in (50, 100):
        for B in batches:
            call = lambda: ddim_infer_sample(model, diff, n=B, img_size=px, device="cuda:0", out_path="/tmp/sweep.png", steps=steps, eta=0.0)
            if B <= 64:
                timed(call)                      # warm pools / packed weights at this shape
            dt = timed(call)
            r = {"sampler": f"ddim{steps}", "px": px, "batch": B, "seconds": dt, "samples_per_s": B / dt, "ms_per_eval": dt / (steps - 1) * 1e3}
            rows.append(r); print(r, flush=True)
    if px == 256:
        dt = timed(lambda: ddpm_infer_sample(model, diff, n=64, img_size=256, device="cuda:0", out_path="/tmp/sweep_ddpm.png"))
        r = {"sampler": "ddpm1000", "px": 256, "batch": 64, "seconds": dt, "samples_per_s": 64 / dt, "ms_per_eval": dt}
        rows.append(r); print(r, flush=True)
    del model
    torch.cuda.empty_cache()
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "config_sweep.json"), "w"), indent=1)
