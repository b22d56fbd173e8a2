#!/usr/bin/env python
"""Key raw metrics of every kernel in an .ncu-rep (ncu -i rep --page raw --csv). Usage: ncu_summary.py rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
if not STALL:
    STALL = [h for h in hdr if "issue_stalled" in h and h.endswith(".pct")]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for w in WANT:
        if w in hdr:
            print(f"  {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    st = []
    for h in STALL:
        try:
            st.append((float(r[hdr.index(h)]), h))
        except ValueError:
            pass
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  stall {h.split('issue_stalled_')[1][:40]:42s} {v:10.3f}")
