#!/bin/bash
# programmatic dependent launch / wgrad side stream: A/B at 64 px (train + DDIM-100) and for DDIM-100 at 256 px
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu --no-c256 > gpurun_out/r2_ab_$name.json 2> gpurun_out/r2_ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_ab_$name.json").read().strip().splitlines()[-1])
    print("low64 $name", round(d["value"]), round(d["ms_per_step"], 3), "ddim", round(d["ddim100"]["value"], 1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e)
PY
}
run base A=1
run nopdl DDPM_B200_PDL=0
run nooverlap DDPM_B200_WGRAD_OVERLAP=0
run nopdl_nooverlap DDPM_B200_PDL=0 DDPM_B200_WGRAD_OVERLAP=0
run base2 A=1
run nopdl2 DDPM_B200_PDL=0
for pdl in 1 0; do
DDPM_B200_PDL=$pdl timeout 300 python - <<PY
import sys, time, torch, contextlib, io
sys.path.insert(0, ".")
from bench import _build_ours
from ddpm_diffusion_model_b200.testing.ddpim_inference import ddim_infer_sample
dev = torch.device("cuda", 0)
m, d, *_ = _build_ours("celeba256", dev)
def call(steps):
    with torch.autocast("cuda", dtype=torch.bfloat16), contextlib.redirect_stdout(io.StringIO()):
        ddim_infer_sample(m, d, n=16, img_size=256, device="cuda:0", out_path="/tmp/x.png", seed=1, steps=steps, eta=0.0)
call(8); torch.cuda.synchronize(); t0 = time.perf_counter(); call(50); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("c256 ddim50 B=16 PDL=$pdl: %.2f ms/eval" % (dt / 49 * 1e3))
PY
done
