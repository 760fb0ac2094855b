#!/bin/bash
# Round-2 (second session) validation on one B200: the fused BatchNorm/pool, BatchNorm/OutConv and AdamW/pack paths.
#   1. tests/test_gpu_parity.py (holds the new kernel tests)      2. the bench line with the new defaults
#   3. the rest of the -m gpu suite                                4. same-box A/B with the three fusions switched off
#   5. ncu launch list of one step                                 6. ncu --set full of the new HBM-bound kernels
mkdir -p gpurun_out
t0=$(date +%s)
lap() { echo "== $1 rc=$2 t=$(( $(date +%s) - t0 ))s"; }
timeout 420 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -25 > gpurun_out/v_parity.txt; lap parity ${PIPESTATUS[0]}
tail -3 gpurun_out/v_parity.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; lap bench $?
timeout 900 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_parity.py --ignore=tests/test_gpu_scripts.py 2>&1 | tail -25 > gpurun_out/v_rest.txt; lap rest ${PIPESTATUS[0]}
tail -3 gpurun_out/v_rest.txt
B200_FUSE_BN_POOL=0 B200_FUSE_BN_OUTCONV=0 B200_ADAMW_PACK=0 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_unfused.json 2> gpurun_out/r02b_bench_unfused.err; lap unfused $?
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_fused.json 2> gpurun_out/r02b_bench_fused.err; lap fused $?
python - <<'PY'
import json
for f in ("r02b_bench", "r02b_bench_unfused", "r02b_bench_fused"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "fwdbwd ms",
              round(d["fwd_bwd_only"]["ms_per_step"], 2), "launches", d["gpu_launches"], "mem", d.get("peak_mem_gb"), "clk", d.get("clocks", {}).get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 python -m pytest tests/test_gpu_scripts.py -q 2>&1 | tail -25 > gpurun_out/v_scripts.txt; lap scripts ${PIPESTATUS[0]}
tail -3 gpurun_out/v_scripts.txt
CMD="python bench.py --profile-steps 1 --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; lap ncu_list $?
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"pool|outconv_fwd|outconv_bwd|adamw_pack" -s 40 -c 30 -f -o gpurun_out/r02b_prof_fused $CMD > gpurun_out/ncu_fused.log 2>&1; lap ncu_fused $?
ls -la gpurun_out/*.ncu-rep 2>/dev/null
