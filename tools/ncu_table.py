#!/usr/bin/env python3
"""One line of key metrics per captured kernel launch:  python tools/ncu_table.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
cols = {"name": "Kernel Name", "us": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
        "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active", "fma%": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lsu%": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "xu%": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smem_wf%": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
        "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "warps/sched": "smsp__warps_active.avg.per_cycle_active"}
ix = {k: (h.index(v) if v in h else None) for k, v in cols.items()}
units = rows[1]
print(" | ".join(cols))
for r in rows[2:]:
    vals = []
    for k in cols:
        i = ix[k]
        v = r[i] if i is not None else "-"
        if k == "name":
            v = v.split("(")[0][-40:]
        elif i is not None:
            try:
                v = f"{float(v.replace(',', '')):.1f}" + (units[i] if k.startswith("dram_") else "")
            except ValueError:
                pass
        vals.append(v)
    print(" | ".join(vals))
