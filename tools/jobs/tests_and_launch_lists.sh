#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_tests5.log
tail -3 gpurun_out/r2_tests5.log
bash tools/jobs/ncu_launch_lists.sh
