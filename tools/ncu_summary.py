#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed):  python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt] [header text]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
out = [sys.argv[3] if len(sys.argv) > 3 else rep, ""]
names = [r[hdr.index('Kernel Name')] for r in data] if 'Kernel Name' in hdr else []
out.append("kernels: " + " | ".join(n[:60] for n in names))
for k in want:
    if k in hdr:
        i = hdr.index(k)
        out.append(f"{k} [{units[i]}] = {', '.join(r[i] for r in data)}")
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        out.append(f"{h} = {', '.join(r[i] for r in data)}")
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2 and sys.argv[2] != '-':
    open(sys.argv[2], "w").write(txt)
print(txt)
