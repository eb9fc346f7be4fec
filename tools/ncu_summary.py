"""Summarise an .ncu-rep (raw page + per-opcode executed instruction mix from the source page).
usage: python tools/ncu_summary.py <rep> <n_output_pixels_per_launch>"""
import csv, io, subprocess, sys
from collections import Counter
rep, npx = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active']
keys += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for k in keys:
    if k in hdr:
        i = hdr.index(k); print(f"{k:90s} {rows[1][i]:>8s} {rows[2][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
iA, iE = hdr.index('Source'), hdr.index('Instructions Executed')
data = [(r[iA].strip(), int(r[iE])) for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith('0x')]
tot = sum(e for _, e in data)
print(f"executed thread-instructions per output pixel: {tot * 32 / npx:.1f}")
c = Counter()
for s, e in data:
    t = s.split(); op = t[1] if t[0].startswith('@') else t[0]
    c[op.split('.')[0]] += e
print("  ".join(f"{op}:{e * 32 / npx:.1f}" for op, e in c.most_common(26)))
