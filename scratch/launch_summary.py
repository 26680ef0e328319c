"""Per-kernel averages of an ncu --csv launch list (gpu__time_duration + dram bytes)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        k = r[hdr.index('Kernel Name')][:60]; m = r[hdr.index('Metric Name')]
        v = float(r[hdr.index('Metric Value')].replace(',', '')); u = r[hdr.index('Metric Unit')]
        a = agg.setdefault(k, {}); b = a.setdefault(m, [0, 0.0, u]); b[0] += 1; b[1] += v
for k, a in agg.items():
    print(k, " ".join(f"{m.split('__')[-1]}={t/n:.1f}{u}(x{n})" for m, (n, t, u) in a.items()))
