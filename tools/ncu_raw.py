#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (raw page): duration, DRAM bytes, issue/pipe utilisation.  usage: ncu_raw.py file.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__inst_executed.sum", "winst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64cyc%")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print(" | ".join(n for _, n in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:44] if n == "kernel" else r[i][:12]) for i, n in idx))
