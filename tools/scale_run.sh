#!/bin/bash
# One 8-GPU lease: the driver's exact bench command at N = 8 (three consecutive runs), 4, 2, 1, configs[3] at N = 8,
# and the N-rank NCCL parity test.  Output lines land in gpurun_out/scale_*.json, stderr next to them.
mkdir -p gpurun_out
port=29500
run() {  # run <tag> <N> <extra bench args...>
  tag=$1; n=$2; shift 2
  port=$((port + 1))
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  fi
  echo "== $tag rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_$tag.json").read().strip().splitlines()[-1])
    print(f"value {d['value']:.1f} {d['unit']} ms/step {d['ms_per_step']:.2f} e2e {d['e2e']['value']:.1f} clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print("no JSON line:", e)
PY
)"
  grep -m2 -E "b200:|b200_device_error|CUDA error" gpurun_out/scale_$tag.err
}
nvidia-smi -L | wc -l
run n8_1 8 --no-cpu-baseline
run n8_2 8 --no-cpu-baseline
run n8_3 8 --no-cpu-baseline
run n4 4 --no-cpu-baseline
run n2 2 --no-cpu-baseline
run n1 1 --no-cpu-baseline
run cfg3_n8 8 --no-cpu-baseline --size 256 --base-ch 128 --batch 8 --seq-len 16
python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -3
