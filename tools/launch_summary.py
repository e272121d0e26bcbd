#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): count, mean, total, share."""
import csv, collections, re, sys

def summarize(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    d = collections.defaultdict(list)
    for x in csv.DictReader(lines):
        if x.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        try:
            v = float(x['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = x['Metric Unit']
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        d[re.sub(r'\(.*', '', x['Kernel Name']).replace('<unnamed>::', '').replace('void ', '')].append(v)
    return d

if __name__ == '__main__':
    d = summarize(sys.argv[1])
    tot = sum(sum(v) for v in d.values())
    print("kernel,launches,mean_us,total_ms,share")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k},{len(v)},{sum(v)/len(v):.1f},{sum(v)/1000:.3f},{sum(v)/tot:.3f}")
