"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.defaultdict(float); n = collections.Counter(); seq = []
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    elif r[ui] == 's': v *= 1e6
    name = r[ki].split('(')[0]
    tot[name] += v; n[name] += 1; seq.append((name, v))
total = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:44s} n={n[k]:5d} total {tot[k]/1e3:9.3f} ms ({100*tot[k]/total:5.1f} %)  avg {tot[k]/n[k]:9.1f} us")
if len(sys.argv) > 2:
    for name, v in seq[:int(sys.argv[2])]: print("  ", name, f"{v:.1f} us")
