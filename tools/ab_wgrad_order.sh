for shape in "4096 1024 4" "2048 512 8" "1024 256 16" "512 512 8" "256 256 16"; do
  for v in 0 1; do echo -n "taps_inner=$v: "; B200_WGRAD_TAPS_INNER=$v python tools/bench_ops.py wgrad $shape 20 256 2>&1 | tail -1; done
done
for v in 0 1; do
  B200_WGRAD_TAPS_INNER=$v ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wgrad -s 2 -c 1 python tools/bench_ops.py wgrad 4096 1024 4 20 256 2>&1 | grep -E "dram__bytes|gpu__time" | sed "s/^/taps_inner=$v /"
done
for v in 0 1 0 1; do echo -n "step taps_inner=$v: "; B200_WGRAD_TAPS_INNER=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d[\"ms_per_step\"], d[\"value\"], d[\"clocks\"][\"sm_mhz\"])"; done
