run() { name=$1; shift; python tools/bench_ops.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s 2 -c 1 -f -o gpurun_out/prof_$name python tools/bench_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1; cat gpurun_out/plain_$name.log; }
KREG=conv run cv64n64 conv 64 64 64 20 256
KREG=conv run cv128n128 conv 128 128 32 20 256
