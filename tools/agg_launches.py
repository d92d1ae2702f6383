"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    k = r[ki].replace('<unnamed>::', '')[:48]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("%-50s n=%4d %10.1f us %5.1f%%  avg %8.1f us" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
print("total us %.0f launches %d" % (tot, sum(a[0] for a in agg.values())))
