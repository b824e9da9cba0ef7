"""Turn the artefacts of an evidence run into the profiles/ documents:
    profile_report.py TAG   reads gpurun_out/{bench_TAG.json, launches_TAG_lane1.csv, prof_TAG_full_raw.csv}
and prints (a) the bench summary, (b) the FULL-step share table (ncu vs live), (c) the ncu --set full table."""
import csv, io, json, os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, 'gpurun_out')
d = json.load(open(os.path.join(G, 'bench_%s.json' % tag)))
r = d['roofline']
print('value %.0f e2e %.0f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
print('sampler %.0f cached %.0f' % (d['apm_iters_per_s']['value'], d['cached_estimates_per_s']))
print('roofline', r['kernel'], 'achieved %.2f peak %.2f frac %.3f traffic %s avg_launch_ms %.3f' % (r['achieved'], r['peak'], r['frac'], r['traffic'], r['avg_launch_ms']))
print('whole step %.1f TF (%.3f), executed %.1f TF' % (r['whole_step']['achieved'], r['whole_step']['frac'], r['whole_step']['executed_tflops']))
for k, v in r['kernels'].items():
    print('   ', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
ref = os.path.join(G, 'bench_%s_reference.json' % tag)
if os.path.isfile(ref):
    rr = json.load(open(ref)); print('reference arm %.1f est/s on %d cores -> x%.0f' % (rr['value'], rr['cpu_baseline']['cores'], d['e2e']['value'] / rr['value']))
print('cpu_baseline', d['cpu_baseline'])
fam = {'k_chol_dataflow': 'k_chol', 'k_chol_step': 'k_chol', 'k_syrk_rev': 'k_syrk_sub', 'k_syrk_lk': 'k_syrk_sub', 'k_symv_lower': 'k_matvec', 'k_symv_reduce': 'k_matvec',
       'k_lt_matvec': 'k_matvec', 'k_l_matvec_rev': 'k_matvec', 'k_fnew_from_s': 'k_matvec', 'k_make_Y': 'k_transpose_u', 'k_antitranspose': 'k_transpose_u',
       'k_transpose_u': 'k_transpose_u', 'k_newton_prep': 'k_newton_vec', 'k_newton_finish': 'k_newton_vec', 'k_is_logw': 'k_is_epilogue'}
lines = [l for l in open(os.path.join(G, 'launches_%s_lane1.csv' % tag)) if not l.startswith('==')]
seq = [(x['Kernel Name'].split('(')[0].replace('apm::', ''), float(x['Metric Value'].replace(',', '')) * (1e3 if x['Metric Unit'] in ('us', 'usecond') else 1))
       for x in csv.DictReader(lines) if x['Metric Name'] == 'gpu__time_duration.sum']
idxs = [i for i, (k, _) in enumerate(seq) if k == 'k_build_K']
seg = seq[idxs[0]:idxs[7]]
tot = OrderedDict()
for k, ns in seg:
    f = fam.get(k, k if k in ('k_build_K', 'k_trsv2', 'k_trsm_rows', 'k_gemm_tri', 'k_is_epilogue') else 'misc')
    e = tot.setdefault(f, [0, 0.]); e[0] += 1; e[1] += ns
T = sum(v[1] for v in tot.values())
live = r['kernels']
print('\n## The 7 FULL steps (%d launches, %.1f ms under ncu)\n' % (len(seg), T / 1e6))
print('| family | launches | ncu us | ncu share | live share (bench.py) |\n|---|---:|---:|---:|---:|')
for k, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('| %s | %d | %.1f | %.1f%% | %s |' % (k, c, ns / 1e3, 100 * ns / T, ('%.1f%%' % (100 * live[k]['share_of_step'])) if k in live else '-'))
rows = list(csv.reader(open(os.path.join(G, 'prof_%s_full_raw.csv' % tag))))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('gpu__time_duration.sum', 'us'), ('launch__registers_per_thread', 'regs'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'dmma%'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64inst%'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('dram__bytes_read.sum', 'dramR'), ('dram__bytes_write.sum', 'dramW'),
        ('lts__t_sector_hit_rate.pct', 'L2hit%'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%')]
cols = [(c, n) for c, n in cols if c in idx]
print('\n| ' + ' | '.join(n for _, n in cols) + ' |\n|' + '---|' * len(cols))
for row in data:
    if len(row) < len(hdr):
        continue
    vals = []
    for c, n in cols:
        v = row[idx[c]]
        if n == 'kernel':
            v = v.split('(')[0].replace('apm::', '')
        else:
            try:
                v = '%.4g' % float(v.replace(',', ''))
            except ValueError:
                pass
        vals.append(v + (' ' + rows[1][idx[c]] if n in ('dramR', 'dramW', 'us') else ''))
    print('| ' + ' | '.join(vals) + ' |')
