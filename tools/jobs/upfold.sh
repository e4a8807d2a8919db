#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_tests6.log
tail -5 gpurun_out/r2_tests6.log
for f in 1 0; do
  DDPM_B200_FOLD_UPSAMPLE=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu > gpurun_out/r2_bench6_fold$f.json 2> gpurun_out/r2_bench6_fold$f.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench6_fold$f.json").read().strip().splitlines()[-1])
    print("fold=$f", d["value"], d["ms_per_step"], d["ddim100"]["value"], d["ddim100"]["ms_per_eval"], d["loss"])
except Exception as e:
    print("fold=$f failed", e)
PY
done
