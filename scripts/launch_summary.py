"""Per-kernel totals of an ncu launch list (csv from --metrics gpu__time_duration.sum): python scripts/launch_summary.py file.csv"""
import collections, csv, statistics, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
d = collections.defaultdict(list)
for row in r:
    v = float(row[vi].replace(',', ''))
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row[ui], 1)
    d[row[ki].split('(')[0]].append(v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-46s n=%5d total %8.2f ms share %.3f median %8.1f us max %8.1f us" % (k[:46], len(v), sum(v) / 1e6, sum(v) / tot, statistics.median(v) / 1e3, max(v) / 1e3))
print("total %.2f ms" % (tot / 1e6))
