"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (usage: file [title])."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:100]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("# total %.0f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
print("share%   time_us  launches  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%6.2f %9.1f %6d  %s' % (100 * v[1] / tot, v[1], v[0], k))
