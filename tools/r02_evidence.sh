#!/bin/bash
# Round-2 evidence run on one B200 (no profiler): smoke, the full -m gpu suite, the bench line, the tf32 mode, the
# per-entry-point breakdown, and the gate-recompute switch at configs[1] and configs[3].
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | grep "^smoke"
python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown --timeline gpurun_out/r02_timeline.csv > gpurun_out/r02_bench_bd.json 2> gpurun_out/r02_breakdown.txt; echo "breakdown rc=$?"
python bench.py --precision tf32 --batch 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_tf32.json 2> gpurun_out/r02_bench_tf32.err; echo "tf32 rc=$?"
for rec in 0 1; do
  B200_GATE_RECOMPUTE=$rec python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_cfg1_rec$rec.json 2> gpurun_out/r02_cfg1_rec$rec.err; echo "cfg1 rec=$rec rc=$?"
  B200_GATE_RECOMPUTE=$rec python bench.py --steps 5 --warmup 3 --no-cpu-baseline --size 256 --base-ch 128 --batch 8 --seq-len 16 > gpurun_out/r02_cfg3_rec$rec.json 2> gpurun_out/r02_cfg3_rec$rec.err; echo "cfg3 rec=$rec rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "mem GB", round(d["peak_mem_gb"], 1), "cell frac", d["roofline"]["frac"] and round(d["roofline"]["frac"], 3),
              "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
    except Exception as e:
        print(f, "ERR", e)
PY
