import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0].replace('void ','')
    v = float(row['Metric Value'].replace(',','')); unit = row['Metric Unit']
    v = v/1e3 if unit == 'ns' else (v*1e3 if unit == 'ms' else v)
    agg.setdefault((name,row.get('Grid Size','')), []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f'{"kernel":44s} {"grid":16s} {"n":>3s} {"mean_us":>9s} {"share":>6s}')
for (n,g),v in agg.items():
    print(f'{n[:44]:44s} {g:16s} {len(v):3d} {sum(v)/len(v):9.1f} {100*sum(v)/tot:5.1f}%')
