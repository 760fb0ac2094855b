# ncu --set full of the two recurrent backward kernels at the temporal cell's shape (Ch 1024 @ 4x4, B 256):
# the whole-sequence weight gradient (K = T*B*H*W) and one BPTT data-gradient step
run() { name=$1; shift; python tools/bench_ops.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s 2 -c 1 -f -o gpurun_out/prof_$name python tools/bench_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1; cat gpurun_out/plain_$name.log; }
KREG=wgrad run wg_temporal wgrad 4096 1024 4 20 256
KREG=conv run dgrad_temporal conv 4096 2048 4 1 256
