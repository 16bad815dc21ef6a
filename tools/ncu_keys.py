#!/usr/bin/env python
"""Key raw metrics of one kernel of an .ncu-rep: tools/ncu_keys.py report.ncu-rep [extra-metric-substring ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, units, r = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__shared_mem_config_size", "launch__shared_mem_per_block_static", "smsp__cycles_active.avg"]
keys += [k for k in h for s in sys.argv[2:] if s in k]
for k in keys:
    if k in h:
        i = h.index(k); print(f"{k:75s} {r[i]:>18s} {units[i]}")
