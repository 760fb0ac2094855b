#!/bin/bash
# Second validation of the round-2 second-session work (after the fixes of the first): parity file under xdist (a crash names
# its test), the rest of the -m gpu suite without the script tests, the bench line, the ncu launch list of one step.
mkdir -p gpurun_out
t0=$(date +%s)
lap() { echo "== $1 rc=$2 t=$(( $(date +%s) - t0 ))s"; }
export PYTHONFAULTHANDLER=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -n 1 --max-worker-restart 4 > gpurun_out/v2_parity.txt 2>&1; lap parity $?
grep -n "crashed\|^FAILED\|passed\|failed" gpurun_out/v2_parity.txt | head -12
timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; lap bench $?
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c_bench.json").read().strip().splitlines()[-1])
    print("bench value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "fwdbwd ms",
          round(d["fwd_bwd_only"]["ms_per_step"], 2), "launches", d["gpu_launches"], "mem", d.get("peak_mem_gb"), "clk", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("bench ERR", e)
PY
timeout 300 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_parity.py --ignore=tests/test_gpu_scripts.py > gpurun_out/v2_rest.txt 2>&1; lap rest $?
tail -3 gpurun_out/v2_rest.txt
CMD="python bench.py --profile-steps 1 --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02c_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; lap ncu_list $?
