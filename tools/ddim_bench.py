"""DDIM sampling sweep on one GPU (BASELINE config 5): samples/s for steps in {50,100}, B in a list,
with and without CUDA graphs (run under gpurun)."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LOW_GPU
from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
from ddpm_diffusion_model_b200.model.unet_backbone import build_unet_64x64
from ddpm_diffusion_model_b200.testing.ddpim_inference import ddim_infer_sample
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_unet_64x64(**LOW_GPU).to(dev).eval()
diff = Diffusion(T=1000, img_size=64).to(dev)
batches = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,8,64,256,512").split(",")]
for graphs in ("0", "1"):
    os.environ["DDPM_B200_GRAPHS"] = graphs
    for steps in (100,):
        for B in batches:
            def call():
                with torch.autocast("cuda", dtype=torch.bfloat16), contextlib.redirect_stdout(io.StringIO()):
                    return ddim_infer_sample(model, diff, n=B, img_size=64, device="cuda:0", out_path="/tmp/ddim_sweep.png", steps=steps, eta=0.0)
            g0 = call(); torch.cuda.synchronize(); t0 = time.perf_counter(); g1 = call(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"graphs={graphs} steps={steps} B={B:4d}: {dt*1e3:8.1f} ms  {B/dt:8.1f} samples/s  {dt/(steps-1)*1e3:6.2f} ms/eval  same_output={bool(torch.equal(g0, g1))}", flush=True)
