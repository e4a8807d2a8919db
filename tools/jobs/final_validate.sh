#!/bin/bash
# what the driver runs at round end, in the same order: GPU tests, smoke(), bench (ours), bench (reference arm)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log; tail -2 gpurun_out/r2_final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
T0=$(date +%s); timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "ours arm wall $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "reference arm wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("train", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "ddim", round(d["ddim100"]["value"], 1), "roof", round(d["roofline"]["frac"], 3),
      "enqueue", round(d["host_enqueue_ms_per_step"], 2), "launches", d["gpu_launches"])
for k, v in (d.get("configs") or {}).items(): print(" ", k, round(v["value"], 2), v["unit"])
print(" vs eager", d.get("vs_gpu_eager"))
print(" cpu", d.get("cpu_baseline"))
r = json.loads(open("gpurun_out/r2_final_ref.json").read().strip().splitlines()[-1])
print("reference arm", round(r["value"], 2), r["unit"], r["cpu_baseline"]["cores"], "cores; ddim", r.get("ddim100", {}).get("value"))
PY
