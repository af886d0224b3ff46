"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]
ki, vi = cols.index('Kernel Name'), cols.index('Metric Value')
agg, tot = collections.OrderedDict(), 0.0
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    name = re.sub(r'\(.*', '', r[ki])[:72]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"total {tot/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg {t/n/1e3:9.1f} us  {k}")
