"""Summarise an .ncu-rep (raw page) into a compact per-launch table.  Usage: ncu_summary.py rep [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('gpu__time_duration.sum', 'us'),
        ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'dmma%'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64pipe%'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64inst%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('dram__bytes_read.sum', 'dramR'), ('dram__bytes_write.sum', 'dramW'),
        ('lts__t_sector_hit_rate.pct', 'L2hit%'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conf'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%')]
cols = [(c, n) for c, n in cols if c in idx]
lines = ['| ' + ' | '.join(n for _, n in cols) + ' |', '|' + '---|' * len(cols)]
for r in data:
    if len(r) < len(hdr):
        continue
    vals = []
    for c, n in cols:
        v = r[idx[c]]
        if n == 'kernel':
            v = v.split('(')[0].replace('apm::', '')
        else:
            try:
                v = '%.4g' % float(v.replace(',', ''))
            except ValueError:
                pass
        vals.append(v + (' ' + rows[1][idx[c]] if n in ('dramR', 'dramW', 'us') else ''))
    lines.append('| ' + ' | '.join(vals) + ' |')
out = '\n'.join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(out + '\n')
