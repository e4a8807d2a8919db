#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/kbench.py --only wgrad --shape "192,192,32" > gpurun_out/r2_plain_wgrad2.log 2>&1 &&
$NCU -k regex:wgrad_tc2_kernel -s 6 -c 1 -o gpurun_out/r2_prof_wgrad_pair_192_32 -f python tools/kbench.py --only wgrad --shape "192,192,32" > gpurun_out/r2_ncu_wgrad2.log 2>&1
python tools/kbench.py --only conv --shape "96,96,64" > gpurun_out/r2_plain_conv.log 2>&1 &&
$NCU -k regex:conv_tc2_kernel -s 6 -c 1 -o gpurun_out/r2_prof_conv_96_96_64 -f python tools/kbench.py --only conv --shape "96,96,64" > gpurun_out/r2_ncu_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
