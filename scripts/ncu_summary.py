"""Summarise an .ncu-rep (raw page + top stalled SASS lines) into a small text file for profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__inst_executed_op_tma_ld.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_selected']
lines = [f'# ncu --set full summary of {rep}', '']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    lines.append(f'## {name}')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f'  {k:78s} {r[i]} {units[i]}')
    lines.append('')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
i = 0
while i < len(srows):
    if srows[i] and srows[i][0] == 'Kernel Name':
        kname = srows[i][1]
        h = srows[i + 1]
        j = i + 2
        body = []
        while j < len(srows) and not (srows[j] and srows[j][0] == 'Kernel Name'):
            body.append(srows[j]); j += 1
        si, ci, ei = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
        items = []
        for k, r in enumerate(body):
            try:
                items.append((float(r[ci]), k, r[ei], r[si].strip()))
            except Exception:
                pass
        tot = sum(v for v, *_ in items) or 1
        lines.append(f'## top stalled SASS instructions: {kname[:110]}  ({int(tot)} samples)')
        for v, k, e, s in sorted(items, reverse=True)[:12]:
            lines.append(f'  {v / tot * 100:5.1f}%  #{k:<5d} exec={e:>9s}  {s[:100]}')
        lines.append('')
        i = j
    else:
        i += 1
open(out, 'w').write('\n'.join(lines))
print('\n'.join(lines[:12]))
