#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "wgrad" 2>&1 | tail -12 > gpurun_out/r2_tests7.log
tail -5 gpurun_out/r2_tests7.log
if grep -q "passed" gpurun_out/r2_tests7.log && ! grep -q "failed" gpurun_out/r2_tests7.log; then
  timeout 300 python tools/kbench.py --only wgrad > gpurun_out/r2_kbench_wgrad_pair.txt 2>&1
  timeout 300 python tools/kbench.py --only wgrad --tcexp 1024 > gpurun_out/r2_kbench_wgrad_single.txt 2>&1
  paste <(grep "^wgrad" gpurun_out/r2_kbench_wgrad_pair.txt | awk '{print $2, $3, $4}') <(grep "^wgrad" gpurun_out/r2_kbench_wgrad_single.txt | awk '{print $3, $4}') | grep -v skip
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
  for e in 0 1024; do
    timeout 600 python bench.py --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu --no-ddim --tc-exp $e > gpurun_out/r2_bench7_exp$e.json 2> gpurun_out/r2_bench7_exp$e.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench7_exp$e.json").read().strip().splitlines()[-1])
    print("tc-exp=$e", d["value"], d["ms_per_step"], d["loss"])
except Exception as ex:
    print("tc-exp=$e failed", ex)
PY
  done
fi
