"""ncu --csv launch list (gpu__time_duration.sum) -> per-kernel table.  Usage: launch_summary.py launches.csv [top]"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki][:66]].append(float(r[vi].replace(',', '')) / (1e3 if r[ui] == 'ns' else 1))
tot = sum(sum(v) for v in d.values()); n = sum(len(v) for v in d.values())
print("# total %.1f ms over %d launches" % (tot / 1e3, n))
print("%-66s %5s %10s %9s %6s" % ("kernel", "n", "sum_us", "max_us", "share"))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:top]:
    print("%-66s %5d %10.1f %9.1f %5.1f%%" % (k, len(v), sum(v), max(v), 100 * sum(v) / tot))
mine = sum(sum(v) for k, v in d.items() if 'scan::' in k)
print("# scan:: kernels %.1f ms (%.1f%%), library kernels %.1f ms" % (mine / 1e3, 100 * mine / tot, (tot - mine) / 1e3))
