#!/bin/bash
# GPU job: ncu --set full captures (one per kernel family under study) -- plain run first, then the same command under ncu
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/kbench.py --only wgrad --shape "192,192,32;96,96,64" > gpurun_out/r2_plain_wgrad.log 2>&1 &&
$NCU -k regex:wgrad_tc_kernel -s 6 -c 1 -o gpurun_out/r2_prof_wgrad_192_32 -f python tools/kbench.py --only wgrad --shape "192,192,32" > gpurun_out/r2_ncu_wgrad.log 2>&1
python tools/kbench.py --only gn --gnshape 96,64 --gnslab 1 > gpurun_out/r2_plain_gn1.log 2>&1 &&
$NCU -k regex:gn_.*slab -s 4 -c 4 -o gpurun_out/r2_prof_gn_slab -f python tools/kbench.py --only gn --gnshape 96,64 --gnslab 1 > gpurun_out/r2_ncu_gn1.log 2>&1
python tools/kbench.py --only gn --gnshape 96,64 --gnslab 0 > gpurun_out/r2_plain_gn0.log 2>&1 &&
$NCU -k regex:gn_.*_kernel -s 10 -c 8 -o gpurun_out/r2_prof_gn_stream -f python tools/kbench.py --only gn --gnshape 96,64 --gnslab 0 > gpurun_out/r2_ncu_gn0.log 2>&1
ls -la gpurun_out/*.ncu-rep
