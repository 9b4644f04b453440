"""profiles/r02_traffic.json from the launch list of tools/profile_r2_final.sh: DRAM bytes per TIMED GROUP of bench.py's kernel timers
(one group = the launches of that kind in one device step: eval_paths = eval_dp_kernel of every class + walk_kernel, scatter =
scatter_kernel + scatter_big_kernel, act_compact = act_count_kernel + act_compact_kernel), second of the two identical batches."""
import collections, csv, io, json, sys
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
L = collections.OrderedDict()
for r in csv.DictReader(io.StringIO(txt)):
    L.setdefault(int(r['ID']), {'name': r['Kernel Name'].split('(')[0].split('<')[0].replace('void ', '')})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
ids = sorted(L)
lev = [i for i in ids if L[i]['name'] == 'lev_match_kernel']
start = lev[1]
group_of = {'eval_dp_kernel': 'eval_dp_kernel', 'walk_kernel': 'eval_dp_kernel', 'scatter_kernel': 'scatter_kernel', 'scatter_big_kernel': 'scatter_kernel',
            'act_compact_kernel': 'act_compact_kernel', 'act_count_kernel': 'act_compact_kernel'}
steps = sum(1 for i in ids if i >= start and L[i]['name'] == 'scatter_kernel')
per = collections.defaultdict(lambda: collections.defaultdict(float))
for i in ids:
    if i < start:
        continue
    x = L[i]
    g = group_of.get(x['name'], x['name'])
    per[g]['launches'] += 1
    per[g]['rd'] += x['dram__bytes_read.sum']
    per[g]['wr'] += x['dram__bytes_write.sum']
    per[g]['us'] += x['gpu__time_duration.sum'] / 1e3
    per[g]['members'] = per[g].get('members', 0)
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/prof_keyword.py "
                 "(DOCS=10000000 VOCAB=1500000 SCORING=detailed, single lane, second of two identical 1024-query batches; tools/profile_r2_final.sh)",
       "unit": "bytes per timed group = per device step (the unit of bench.py's roofline.launches / algorithmic_bytes_per_launch)",
       "device_steps_with_activations": steps, "kernels": {}}
for g, v in per.items():
    n = steps if g in ('eval_dp_kernel', 'scatter_kernel', 'act_compact_kernel') else v['launches']
    out["kernels"][g] = {"groups": n, "kernel_launches": int(v['launches']), "dram_bytes_per_launch": (v['rd'] + v['wr']) / n,
                         "dram_read_per_launch": v['rd'] / n, "dram_write_per_launch": v['wr'] / n, "ncu_us_per_batch": round(v['us'], 1)}
json.dump(out, sys.stdout, indent=1)
