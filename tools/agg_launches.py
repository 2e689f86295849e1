"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d['Metric Value'].replace(',', ''))
            except ValueError:
                continue
            unit = d.get('Metric Unit', 'ns')
            v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'ns ': 1.0}.get(unit, 1.0)
            agg[d['Kernel Name'][:70]][0] += 1
            agg[d['Kernel Name'][:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"== {path}: {tot / 1e6:.3f} ms total")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:70s} n={v[0]:4d} tot={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e3:9.1f} us {100 * v[1] / tot:5.1f}%")
