#!/bin/bash
# GPU job: TMA slab GroupNorm kernels -- parity, kbench A/B (streaming vs slab vs slab without 16-CTA clusters), step A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_tests4.log
tail -6 gpurun_out/r2_tests4.log
timeout 600 python tools/kbench.py --only gn --gnslab 0,1,2 > gpurun_out/r2_kbench_gn_tma.txt 2>&1
grep -E "per train step|---" gpurun_out/r2_kbench_gn_tma.txt
for m in 0 1 2; do
  if [ $m = 2 ]; then export DDPM_B200_GN_SLAB=1 DDPM_B200_GN_SLAB_CS16=0; else export DDPM_B200_GN_SLAB=$m DDPM_B200_GN_SLAB_CS16=1; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu > gpurun_out/r2_bench4_slab$m.json 2> gpurun_out/r2_bench4_slab$m.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench4_slab$m.json").read().strip().splitlines()[-1])
    print($m, d["value"], d["ms_per_step"], d["ddim100"]["value"], d["host_enqueue_ms_per_step"], d["loss"])
except Exception as e:
    print($m, "failed", e)
PY
done
