#!/bin/bash
# GPU job: ncu launch lists (gpu__time_duration per launch, cold-cache/serialised: compare SHARES) of one train step of
# both bench configs.  The plain run goes first (B200_PROFILING.md: profile only what has exited 0 without ncu).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for cfg in low64 celeba256; do
  python bench.py --profile --no-ddim --no-cpu --config $cfg > gpurun_out/r2_plain_$cfg.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/r2_launches_step_$cfg.csv python bench.py --profile --no-ddim --no-cpu --config $cfg > gpurun_out/r2_ncu_$cfg.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_launches_step_$cfg.csv last-step > gpurun_out/r2_launches_step_${cfg}_summary.txt 2>&1
  head -32 gpurun_out/r2_launches_step_${cfg}_summary.txt
done
