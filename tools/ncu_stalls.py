#!/usr/bin/env python
"""Print the warp-stall breakdown and a few headline metrics of an ncu report."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; vals = rows[2] if len(rows) > 2 else rows[1]
d = dict(zip(hdr, vals))
items = []
for k, v in d.items():
    if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued'):
        try: items.append((int(v), k.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        except ValueError: pass
tot = sum(i[0] for i in items)
for v, k in sorted(items, reverse=True)[:10]: print('%-28s %5.1f%%' % (k, 100 * v / tot))
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__block_size',
          'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']:
    for kk in d:
        if kk == k: print(kk, '=', d[kk])
