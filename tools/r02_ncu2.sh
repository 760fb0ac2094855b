#!/bin/bash
# ncu --set full of the fused cell forward (temporal cell, first of the three per step) and of the pair weight-gradient
# kernel inside one training step at configs[1]
mkdir -p gpurun_out
CMD="python bench.py --profile-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_tc2_kernel<.int.256, .int.1>" -s 3 -c 1 -f -o gpurun_out/r02_prof_cell $CMD > gpurun_out/ncu_cell.log 2>&1
echo "cell rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wgrad_tc2_kernel<.int.256" -s 20 -c 2 -f -o gpurun_out/r02_prof_wgrad2 $CMD > gpurun_out/ncu_wgrad2.log 2>&1
echo "wgrad2 rc=$?"
ls -la gpurun_out/*.ncu-rep
