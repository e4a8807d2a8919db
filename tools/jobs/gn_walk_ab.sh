#!/bin/bash
# GPU job: GroupNorm streaming kernels with incremental addressing (PixWalk) -- parity, then same-box A/B of builds
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=ddpm-diffusion-model_b200/lib
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests3.log
tail -4 gpurun_out/r2_tests3.log
for v in oldgn "" bwd3 fwd3bwd3; do
  lib=$L/libddpm_b200${v:+_$v}.so
  [ -f $lib ] || continue
  echo "=== build ${v:-current}" | tee -a gpurun_out/r2_kbench_gn_walk.txt
  DDPM_B200_LIB=$PWD/$lib python tools/kbench.py --only gn --gnslab 0 2>&1 | grep -E "FUSED|bwd|per train step" >> gpurun_out/r2_kbench_gn_walk.txt
done
for v in oldgn "" bwd3; do
  lib=$L/libddpm_b200${v:+_$v}.so
  [ -f $lib ] || continue
  DDPM_B200_GN_SLAB=0 DDPM_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu > gpurun_out/r2_bench3_${v:-current}.json 2> gpurun_out/r2_bench3_${v:-current}.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench3_${v:-current}.json").read().strip().splitlines()[-1])
print("${v:-current}", d["value"], d["ms_per_step"], d["ddim100"]["value"], d["host_enqueue_ms_per_step"])
PY
done
grep -E "per train step|===" gpurun_out/r2_kbench_gn_walk.txt
