"""Top source lines / stall reasons of one launch in an .ncu-rep.  Usage: ncu_source_top.py rep launch_index [regex]"""
import csv, subprocess, sys, io, collections
rep, li = sys.argv[1], int(sys.argv[2])
rx = sys.argv[3] if len(sys.argv) > 3 else '.'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx,
                      '--launch-skip', str(li), '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
his = [i for i, r in enumerate(rows) if r and r[0] in ('Address', 'Line')]
print('tables:', [(i, rows[i][:2]) for i in his][:4])
for hi in his[:2]:
    hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
    if '# Samples' not in idx: continue
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = collections.Counter(); n = 0
    def num(x):
        try: return int(x)
        except: return 0
    for r in data:
        n += num(r[idx['# Samples']])
        for s in stalls: tot[s] += num(r[idx[s]])
    print('=== table keyed by', hdr[0], 'rows', len(data), 'samples', n)
    print('  '.join('%s %.1f%%' % (s.replace('stall_', ''), 100. * v / max(n, 1)) for s, v in tot.most_common(9)))
    top = sorted(data, key=lambda r: -num(r[idx['# Samples']]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 22]
    for r in top:
        st = {s.replace('stall_', ''): num(r[idx[s]]) for s in stalls if num(r[idx[s]]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print('%6s %5.1f%%  %-110s %s' % (r[idx['# Samples']], 100. * num(r[idx['# Samples']]) / max(n, 1), r[idx['Source']][:110], st))
