# usage: bash tools/prof_ops.sh  -- ncu --set full on single kernels of tools/bench_ops.py (developer aid)
run() { name=$1; shift; python tools/bench_ops.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s 2 -c 1 -f -o gpurun_out/prof_$name python tools/bench_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1; cat gpurun_out/plain_$name.log; }
KREG=wgrad run wgnarrow2 wgrad 64 64 64 5 128
KREG=conv run cv128n64 conv 128 64 64 5 128
KREG=conv run cv64n64b conv 64 64 64 5 128
KREG=wgrad run wg128 wgrad 128 128 32 5 256
