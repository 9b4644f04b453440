"""Summarise a tools/profile_r2_list.sh capture: per-kernel totals of the second batch and the heaviest eval launches."""
import collections, csv, io, sys
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
L = collections.OrderedDict()
for r in csv.DictReader(io.StringIO(txt)):
    k = int(r['ID'])
    L.setdefault(k, {'name': r['Kernel Name'].split('(')[0].replace('void ', ''), 'grid': int(r['Grid Size'].strip('()').split(',')[0])})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
ids = sorted(L)
lev = [i for i in ids if L[i]['name'].startswith('lev_match')]
start = lev[1] if len(lev) > 1 else 0
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0])
for i in ids:
    if i < start:
        continue
    x = L[i]
    n = x['name']
    if n.startswith('eval_dp'):
        n += '/smem%d' % int(x.get('launch__shared_mem_per_block_dynamic', 0))
    a = agg[n]
    a[0] += 1; a[1] += x['gpu__time_duration.sum'] / 1e3; a[2] += x['dram__bytes_read.sum']; a[3] += x['dram__bytes_write.sum']; a[4] += x['smsp__inst_executed.sum']; a[5] += x['grid']
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:36s} n={a[0]:4d} t={a[1]:9.1f}us {100*a[1]/tot:5.1f}% rd={a[2]/1e6:8.1f}MB wr={a[3]/1e6:8.1f}MB inst={a[4]/1e6:8.1f}M ctas={a[5]} inst/cta={a[4]/max(1,a[5]):.0f}")
print('total us', round(tot, 1))
pat = sys.argv[2] if len(sys.argv) > 2 else 'eval_dp'
ev = [(L[i]['gpu__time_duration.sum'] / 1e3, i) for i in ids if i >= start and L[i]['name'].startswith(pat)]
for t, i in sorted(ev, reverse=True)[:14]:
    x = L[i]
    print(i, x['grid'], int(x.get('launch__shared_mem_per_block_dynamic', 0)), f"{t:.1f}us rd={x['dram__bytes_read.sum']/1e6:.1f}MB wr={x['dram__bytes_write.sum']/1e6:.1f}MB occ={x['sm__warps_active.avg.pct_of_peak_sustained_active']:.0f}% inst={x['smsp__inst_executed.sum']/1e6:.1f}M inst/cta={x['smsp__inst_executed.sum']/x['grid']:.0f}")
