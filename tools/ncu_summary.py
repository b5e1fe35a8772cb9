#!/usr/bin/env python
"""Print the handful of ncu metrics we track from a .ncu-rep (one line per captured launch)."""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    for r in rows[2:]:
        print("==", rep, r[idx["Kernel Name"]][:90])
        for k in KEYS:
            if k in idx:
                print(f"   {k:75s} {r[idx[k]]:>16s} {U[idx[k]]}")
        st = sorted(((float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in idx.items() if h.startswith(STALL) and h.endswith("per_issue_active.ratio")), reverse=True)
        print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
