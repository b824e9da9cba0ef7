"""Per-source-line stall samples of one kernel: joins the SASS addresses of an .ncu-rep source page with the line table of
the binary (nvdisasm -g).  Usage: ncu_line_profile.py <rep> <binary-or-so> <kernel-substring> [top] [launch]
(launch: index of the launch inside a report that holds several, default 0)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, binary, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
launch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(binary)], cwd=tmp, capture_output=True)
dis = ''
for f in os.listdir(tmp):
    if f.endswith('.cubin'):
        dis += subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
# offset -> (file, line) inside the kernel's section
off2line, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith('//--------------------- .text.'):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[launch]
end = his[launch + 1] - 1 if launch + 1 < len(his) else len(rows)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) and r[0] != 'Address']
num = lambda x: int(x) if x.isdigit() else 0
base = min(int(r[idx['Address']], 16) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
per = collections.defaultdict(collections.Counter)
tot = 0
for r in data:
    off = int(r[idx['Address']], 16) - base
    key = off2line.get(off, ('?', 0))
    s = num(r[idx['# Samples']])
    tot += s
    per[key]['samples'] += s
    for st in stalls:
        per[key][st.replace('stall_', '')] += num(r[idx[st]])
print('total samples', tot)
src_cache = {}
def src(f, l):
    for root in ('auxiliary-pm-mcmc_b200/csrc', 'scripts'):
        p = os.path.join(root, f)
        if os.path.isfile(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if 0 < l <= len(src_cache[p]) else ''
    return ''
for key, c in sorted(per.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    st = sorted(((k, v) for k, v in c.items() if k != 'samples' and v), key=lambda kv: -kv[1])[:3]
    print('%5.1f%%  %s:%d  %-90s %s' % (100. * c['samples'] / tot, key[0], key[1], src(*key), dict(st)))
