#!/bin/bash
# Prints the key roofline metrics of every kernel in an .ncu-rep (read here, no GPU needed).
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = {h: i for i, h in enumerate(hdr)}
for d in data:
    print("-" * 100)
    for w in want:
        if w in idx:
            print(f"  {w:75s} {d[idx[w]][:110]:>20s} {units[idx[w]]}")
'
