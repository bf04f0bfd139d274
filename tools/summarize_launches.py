import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
d = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        d[row['Kernel Name'][:70]].append(v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):3d} avg={sum(v)/len(v):8.2f} us  share={sum(v)/tot*100:5.1f}%")
