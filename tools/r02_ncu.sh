#!/bin/bash
# Round-2 profiler evidence (one gpurun call, after the same commands ran clean without ncu):
#   1. launch list of one training step at configs[1] (gpu__time_duration per launch: compare SHARES)
#   2. ncu --set full of the fused cell forward (temporal cell, T = 20) inside that step, of the pair weight-gradient kernel
#      and of the 1-CTA weight-gradient kernel whose producer ring was fixed this round
mkdir -p gpurun_out
CMD="python bench.py --profile-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_prof.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_tc2_kernel<.int.256, .int.1>" -s 3 -c 1 -f -o gpurun_out/r02_prof_cell $CMD > gpurun_out/ncu_cell.log 2>&1
echo "cell rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wgrad_tc2_kernel<.int.256" -s 20 -c 2 -f -o gpurun_out/r02_prof_wgrad2 $CMD > gpurun_out/ncu_wgrad2.log 2>&1
echo "wgrad2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 4 -c 2 -f -o gpurun_out/r02_prof_wgrad1 $CMD > gpurun_out/ncu_wgrad1.log 2>&1
echo "wgrad1 rc=$?"
ls -la gpurun_out/*.ncu-rep
