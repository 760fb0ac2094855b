#!/bin/bash
# tools/scale_probe.sh <N> [bench args]: one torchrun bench at N ranks, prints the per-step / per-rank profile
n=$1; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/probe_n$n.json 2> gpurun_out/probe_n$n.err
echo rc=$?
python - <<PY
import json
d = json.loads(open("gpurun_out/probe_n$n.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "fb", d["fwd_bwd_only"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
for k, v in d["step_profile"].items():
    print(k, "median/rank", v["per_rank_median_ms"]); print(k, "max/rank", v["per_rank_max_ms"]); print(k, "rank0 steps", v["rank0_steps_ms"])
PY
