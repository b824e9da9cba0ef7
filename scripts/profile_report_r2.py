"""profiles/ documents of a round-2 evidence run (scripts/evidence.sh TAG):
    profile_report_r2.py TAG  ->  profiles/{bench_TAG_b256.json, bench_TAG_reference.json, TAG_launch_summary.md,
                                            TAG_ncu_k_chol_flow_full.md} and the traffic record profiles/ncu_traffic.json"""
import csv, json, os, shutil, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
line = lambda f: json.loads(open(f).read().strip().splitlines()[-1])
d = line(os.path.join(G, 'bench_%s.json' % tag))
shutil.copy(os.path.join(G, 'bench_%s.json' % tag), os.path.join(P, 'bench_%s_b256.json' % tag))
ref = None
if os.path.isfile(os.path.join(G, 'bench_%s_reference.json' % tag)):
    ref = line(os.path.join(G, 'bench_%s_reference.json' % tag))
    shutil.copy(os.path.join(G, 'bench_%s_reference.json' % tag), os.path.join(P, 'bench_%s_reference.json' % tag))
short = line(os.path.join(G, 'short_%s.json' % tag))
CMD = 'python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --apm-iters 2'

# ---- launch list
def fam(k):
    if k.startswith('k_chol_flow_init'): return 'misc'
    if 'k_chol_flow' in k: return 'k_chol'
    for f, names in (('k_matvec', ('k_symv_lower', 'k_symv_reduce', 'k_lt_matvec', 'k_l_matvec_rev', 'k_fnew_from_s')),
                     ('k_transpose_u', ('k_transpose_u', 'k_antitranspose')), ('k_newton_vec', ('k_newton_prep', 'k_newton_finish')),
                     ('k_is_epilogue', ('k_is_logw', 'k_is_epilogue'))):
        if any(k.startswith(n) for n in names): return f
    for f in ('k_build_K', 'k_trsv2', 'k_trsm_rows', 'k_gemm_tri'):
        if k.startswith(f): return f
    if k.startswith('k_sampler'): return 'sampler (normals / ellipse / copies)'
    return 'misc'
lines = [l for l in open(os.path.join(G, 'launches_%s.csv' % tag)) if not l.startswith('==')]
seq = []
for x in csv.DictReader(lines):
    if x['Metric Name'] != 'gpu__time_duration.sum': continue
    v = float(x['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1., 'usecond': 1., 'ms': 1e3, 'msecond': 1e3}[x['Metric Unit']]
    name = x['Kernel Name'].split('(')[0].replace('apm::', '').replace('void ', '')
    seq.append((name, v, 'apm::' in x['Kernel Name'] or name.startswith('k_')))
tot = OrderedDict()
for k, us, ours in seq:
    e = tot.setdefault(k if ours else 'torch (input generation / copies)', [0, 0.]); e[0] += 1; e[1] += us
T = sum(v[1] for v in tot.values())
out = ['# Round 2 (%s): ncu launch list of the short bench command' % tag, '', 'Command: `%s` under `ncu --metrics gpu__time_duration.sum --clock-control none -c 1500`' % CMD,
       '(after the same command had exited 0 without ncu; cold-cache, serialised per-launch times -- compare SHARES, not absolutes).', '',
       '| kernel | launches | total us | share |', '|---|---:|---:|---:|']
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append('| %s | %d | %.1f | %.1f%% |' % (k, c, us, 100 * us / T))
idx = [i for i, (k, _, _) in enumerate(seq) if k == 'k_build_K']
n_steps = 7          # 1 warm-up + 2 host-buffer warm-ups... the first seven k_build_K launches start FULL steps of the 256-chain engine
seg = seq[idx[0]:idx[n_steps]] if len(idx) > n_steps else seq[idx[0]:]
ft = OrderedDict()
for k, us, _ in seg:
    e = ft.setdefault(fam(k), [0, 0.]); e[0] += 1; e[1] += us
TS = sum(v[1] for v in ft.values())
live = short['roofline']['kernels']
out += ['', '## The first %d FULL steps (%d launches, %.1f ms under ncu): family shares next to bench.py\'s live CUDA-event shares of the same command' % (n_steps, len(seg), TS / 1e3), '',
        '| family | launches | ncu us | ncu share | live share (roofline pass of the same command, without ncu) |', '|---|---:|---:|---:|---:|']
for k, (c, us) in sorted(ft.items(), key=lambda kv: -kv[1][1]):
    out.append('| %s | %d | %.1f | %.1f%% | %s |' % (k, c, us, 100 * us / TS, ('%.1f%%' % (100 * live[k]['share_of_step'])) if k in live else '-'))
a, b = idx[1], idx[2]
out += ['', '## One FULL step, launch by launch (second step)', '', '```']
out += ['%-28s %9.1f us' % (k, us) for k, us, _ in seq[a:b]]
out += ['```', 'sum %.1f us' % sum(us for _, us, _ in seq[a:b])]
open(os.path.join(P, '%s_launch_summary.md' % tag), 'w').write('\n'.join(out) + '\n')

# ---- ncu --set full of the ten k_chol_flow launches of one step
rows = list(csv.reader(open(os.path.join(G, 'prof_%s_full_raw.csv' % tag))))
hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if len(r) >= len(rows[0])]
ix = {h: i for i, h in enumerate(hdr)}
cols = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('gpu__time_duration.sum', 'time'), ('launch__registers_per_thread', 'regs'),
        ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'dmma pipe %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'smem ld conflicts')]
cols = [(c, n) for c, n in cols if c in ix]
def cell(r, c, n):
    v = r[ix[c]]
    if n == 'kernel':
        return v.split('(')[0].replace('apm::', '').replace('void ', '')
    try:
        v = '%.4g' % float(v.replace(',', ''))
    except ValueError:
        pass
    return v + (' ' + units[ix[c]] if n in ('time', 'dram read', 'dram write') else '')
o2 = ['# Round 2 (%s): `ncu --set full` of the ten `k_chol_flow` launches of one FULL step' % tag, '',
      'Command: `ncu --set full --clock-control none --import-source on -k regex:^k_chol_flow$ -s 10 -c 10 %s` (after the same command had exited 0 without ncu).' % CMD,
      'Launch order of a step: chol K | B-space rounds 1-3 | round 4: B-space (empty: every chain was predicted to finish), M-space | round 5: B-space (empty), M-space of the ~7 % stragglers | ... (empty launches of rounds whose mask is empty).',
      '`<0, 0>` plain factorisation, `<0, 1>` + fused forward substitution (Newton round), `<1, 1>` + M\' accumulated from L_K and the anti-transposed second store.', '',
      '| ' + ' | '.join(n for _, n in cols) + ' |', '|' + '---|' * len(cols)]
traffic = {}
for r in data:
    o2.append('| ' + ' | '.join(cell(r, c, n) for c, n in cols) + ' |')
    name = cell(r, 'Kernel Name', 'kernel')
    ms = float(r[ix['gpu__time_duration.sum']].replace(',', '')) * {'ms': 1., 'msecond': 1., 'us': 1e-3, 'usecond': 1e-3}[units[ix['gpu__time_duration.sum']]]
    sc = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.}
    rd = float(r[ix['dram__bytes_read.sum']].replace(',', '')) * sc[units[ix['dram__bytes_read.sum']]]
    wr = float(r[ix['dram__bytes_write.sum']].replace(',', '')) * sc[units[ix['dram__bytes_write.sum']]]
    if ms > 1.0:
        traffic.setdefault(name, []).append((ms, rd, wr))
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
big = [r for r in data if '<0, 1' in r[ix['Kernel Name']] and float(r[ix['gpu__time_duration.sum']].replace(',', '')) > 1.0]
if big and stall:
    r = big[0]
    o2 += ['', '## Warp stall reasons of a full-batch Newton-round launch (`<0, 1>`), warps per issue-active cycle', '', '| reason | ratio |', '|---|---:|']
    for h in sorted(stall, key=lambda h: -float(r[ix[h]].replace(',', '') or 0))[:8]:
        o2.append('| %s | %.2f |' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(r[ix[h]].replace(',', ''))))
open(os.path.join(P, '%s_ncu_k_chol_flow_full.md' % tag), 'w').write('\n'.join(o2) + '\n')

# ---- traffic record for bench.py (dominant kernel = the family; the record is the Newton-round launch, the most frequent one)
rec = {}
for name, v in traffic.items():
    ms, rd, wr = v[0]
    rec[name] = {'ms': ms, 'dram_read': rd, 'dram_write': wr}
tf = os.path.join(P, 'ncu_traffic.json')
allrec = json.load(open(tf)) if os.path.isfile(tf) else {}
key = [k for k in rec if '<0, 1' in k]
if key:
    x = rec[key[0]]
    allrec['k_chol:n=768:chains=256'] = {
        'dram_bytes_per_launch': x['dram_read'] + x['dram_write'], 'dram_read': x['dram_read'], 'dram_write': x['dram_write'],
        'source': 'profiles/%s_ncu_k_chol_flow_full.md: ncu --set full of one full-batch Newton-round launch k_chol_flow<0, 1> (256 chains, n=768, '
                  'chol(I+SKS) + inverse blocks + fused forward substitution) inside the bench command; dram__bytes_read.sum + dram__bytes_write.sum; '
                  'other launches of the step: %s' % (tag, json.dumps({k: {'ms': round(v['ms'], 3), 'GB': round((v['dram_read'] + v['dram_write']) / 1e9, 2)} for k, v in rec.items()}))}
    json.dump(allrec, open(tf, 'w'), indent=1)
print(open(os.path.join(P, '%s_ncu_k_chol_flow_full.md' % tag)).read())
r = d['roofline']
print('value %.0f e2e %.0f ms %.3f apm %.0f (py %.0f) frac %.3f plain %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['apm_iters_per_s']['value'],
      d['apm_iters_per_s_python_scheduler']['value'], r['frac'], r['plain_full_batch_launch']['frac']))
if ref: print('reference arm %.1f est/s on %d cores; sampler %.1f it/s; pmmh %.1f' % (ref['value'], ref['cpu_baseline']['cores'], ref['apm_iters_per_s']['value'], ref['configs']['pmmh']['value']))
