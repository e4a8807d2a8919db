#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "groupnorm or celeba256" 2>&1 | tail -4
for cs in 1 0 1 0; do
  DDPM_B200_GN_CS16=$cs timeout 100 python bench.py --config celeba256 --steps 8 --warmup 3 --no-eager --no-cpu --no-c256 > gpurun_out/r2_cs16_$cs.json 2> gpurun_out/r2_cs16.err
  rc=$?; v=$(python -c "import json;d=json.loads(open('gpurun_out/r2_cs16_$cs.json').read().strip().splitlines()[-1]);print(round(d['value'],1), round(d['ms_per_step'],2))" 2>/dev/null)
  echo "c256 train GN_CS16=$cs rc=$rc $v"
done
for cs in 1 0; do
DDPM_B200_GN_CS16=$cs timeout 300 python - <<PY
import sys, time, torch, contextlib, io
sys.path.insert(0, ".")
from bench import _build_ours
from ddpm_diffusion_model_b200.testing.ddpim_inference import ddim_infer_sample
dev = torch.device("cuda", 0)
m, d, *_ = _build_ours("celeba256", dev)
def call(steps, n):
    with torch.autocast("cuda", dtype=torch.bfloat16), contextlib.redirect_stdout(io.StringIO()):
        ddim_infer_sample(m, d, n=n, img_size=256, device="cuda:0", out_path="/tmp/x.png", seed=1, steps=steps, eta=0.0)
for n in (16, 64):
    call(6, n); torch.cuda.synchronize(); t0 = time.perf_counter(); call(30, n); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("c256 ddim B=%d GN_CS16=$cs: %.2f ms/eval" % (n, dt / 29 * 1e3))
PY
done
