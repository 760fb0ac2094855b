set -x
run() { name=$1; shift; python tools/bench_ops.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s 2 -c 1 -f -o gpurun_out/prof_$name python tools/bench_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1; cat gpurun_out/plain_$name.log; }
KREG=wgrad run wgbig wgrad 4096 1024 4 20 256
KREG=conv run dgbig conv 4096 2048 4 1 256
KREG=wgrad run wgnarrow wgrad 64 64 64 5 128
KREG=conv run cvnarrow conv 64 64 64 5 128
