#!/usr/bin/env python3
"""tools/ncu_stalls.py <raw.csv> — warp stall-reason shares and headline metrics from
`ncu -i X.ncu-rep --page raw --csv` output (development aid; summaries go to profiles/)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for data in rows[2:]:
    d = dict(zip(hdr, data))
    print("##", d.get("Kernel Name", "?")[:100], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    st = {}
    for h, v in d.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                st[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(v)
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    print("   stalls:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__cycles_elapsed.max"]
    for k in keys:
        if k in d:
            print(f"   {k} = {d[k]} {units[hdr.index(k)]}")
