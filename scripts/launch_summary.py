"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: launch_summary.py launches.csv "title" "command" [out.md]"""
import csv
import sys
from collections import OrderedDict

path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(path, newline='') as fh:
    lines = [ln for ln in fh if not ln.startswith('==')]
rd = csv.DictReader(lines)
tot = OrderedDict()
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name'].split('(')[0]
    if 'apm::' not in name:
        name = 'torch (input generation / copies)'
    ns = float(r['Metric Value'].replace(',', ''))
    if r.get('Metric Unit') in ('us', 'usecond'):
        ns *= 1e3
    e = tot.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += ns
total = sum(v[1] for v in tot.values())
out = ['# %s' % title, '', 'Command: `%s`' % command,
       '(cold-cache, serialised per-launch times -- compare SHARES, not absolutes)', '',
       '| kernel | launches | total us | share |', '|---|---:|---:|---:|']
for name, (cnt, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append('| %s | %d | %.1f | %.1f%% |' % (name, cnt, ns / 1e3, 100. * ns / total))
text = '\n'.join(out) + '\n'
print(text)
if len(sys.argv) > 4:
    open(sys.argv[4], 'w').write(text)
