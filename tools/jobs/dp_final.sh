#!/bin/bash
# final multi-GPU check: the driver's own command line (full default bench) under torchrun
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_dp$N.json 2> gpurun_out/r2_final_bench_dp$N.err
echo "rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_final_bench_dp$N.json").read().strip().splitlines()[-1])
print("N=$N train", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "ddim", round(d["ddim100"]["value"], 1))
for k, v in (d.get("configs") or {}).items(): print(" ", k, round(v["value"], 2), v["unit"])
PY
T0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_ref_dp$N.json 2> gpurun_out/r2_final_ref_dp$N.err
echo "reference arm rc=$? wall $(( $(date +%s) - T0 )) s"; tail -c 400 gpurun_out/r2_final_ref_dp$N.json
