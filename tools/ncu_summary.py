#!/usr/bin/env python
"""Summarise an .ncu-rep (first profiled launch) into a text file for profiles/.  usage: tools_ncu_summary.py rep out title"""
import csv, subprocess, sys, io
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_atom.sum','lts__t_sectors_srcunit_tex_op_red.sum',
        'smsp__warps_eligible.avg.per_cycle_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
with open(out, "w") as f:
    f.write(title + "\n\n")
    for h, u, v in zip(hdr, units, vals):
        if h in keep or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            f.write(f"{h:90s} {v:>22s} {u}\n")
print(open(out).read())
