"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel for the last md step."""
import csv, re, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_r01.csv'
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
order = []
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']); v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e6 if unit == 'ns' else v / 1e3 if unit == 'us' else v * 1e3 if unit == 's' else v
    order.append((name, v))
idx = [i for i, (n, v) in enumerate(order) if 'k_md_first_half' in n]
seg = order[idx[-1]:] if idx else order
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seg:
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v for n, v in seg)
print('last md step: total kernel ms', round(tot, 3), 'launches', len(seg))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
    print(f'  {n:45s} n={c:4d} total={t:9.3f} ms avg={t/c:8.4f} share={t/tot:6.1%}')
