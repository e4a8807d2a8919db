#!/bin/bash
# GPU job: slab GroupNorm kernels -- parity tests, per-kernel A/B (kbench) and step-level A/B (bench.py)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log
tail -4 gpurun_out/r2_tests2.log
python tools/kbench.py --only gn --gnslab 0,1,2 > gpurun_out/r2_kbench_gn_slab.txt 2>&1
tail -3 gpurun_out/r2_kbench_gn_slab.txt
for m in 0 1; do
  DDPM_B200_GN_SLAB=$m python bench.py --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu > gpurun_out/r2_bench2_slab$m.json 2> gpurun_out/r2_bench2_slab$m.err
done
DDPM_B200_GN_SLAB_CS16=0 python bench.py --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu > gpurun_out/r2_bench2_slab2.json 2> gpurun_out/r2_bench2_slab2.err
python - <<'PY'
import json
for m in (0, 1, 2):
    try:
        d = json.loads(open("gpurun_out/r2_bench2_slab%d.json" % m).read().strip().splitlines()[-1])
        print(m, d["value"], d["ms_per_step"], d["ddim100"]["value"], d["host_enqueue_ms_per_step"])
    except Exception as e:
        print(m, "failed", e)
PY
