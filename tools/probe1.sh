#!/bin/bash
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/probe_n1.json 2> gpurun_out/probe_n1.err; echo rc=$?
python - <<PY
import json
d = json.loads(open("gpurun_out/probe_n1.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "fb", d["fwd_bwd_only"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
for k, v in d["step_profile"].items():
    for kk, vv in v.items(): print(k, kk, vv)
PY
