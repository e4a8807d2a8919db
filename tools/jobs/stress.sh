#!/bin/bash
# stability of the new defaults (PDL off, pair wgrad on): repeated bench runs, both configs, with and without the side stream
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
n=0
for i in 1 2 3 4 5; do
  for ov in 1 0; do
    DDPM_B200_WGRAD_OVERLAP=$ov timeout 100 python bench.py --config celeba256 --steps 8 --warmup 3 --no-eager --no-cpu --no-c256 > gpurun_out/r2_stress_c256_$ov_$i.json 2> gpurun_out/r2_stress.err
    rc=$?; v=$(python -c "import json;d=json.loads(open('gpurun_out/r2_stress_c256_$ov_$i.json').read().strip().splitlines()[-1]);print(round(d['value'],1))" 2>/dev/null)
    echo "c256 overlap=$ov run $i rc=$rc $v"
  done
done
for i in 1 2 3; do
  timeout 100 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu --no-c256 > gpurun_out/r2_stress_low_$i.json 2> gpurun_out/r2_stress.err
  rc=$?; v=$(python -c "import json;d=json.loads(open('gpurun_out/r2_stress_low_$i.json').read().strip().splitlines()[-1]);print(round(d['value']), round(d['ddim100']['value'],1))" 2>/dev/null)
  echo "low64 run $i rc=$rc $v"
done
