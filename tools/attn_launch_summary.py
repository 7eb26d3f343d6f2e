import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki][:60]].append(float(r[vi].replace(',', '')) / (1e3 if r[ui] == 'ns' else 1))
for k, v in d.items():
    print("%-62s n=%3d avg=%9.1f us" % (k, len(v), sum(v) / len(v)))
