"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
kn, mv = h.index('Kernel Name'), h.index('Metric Value')
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    try:
        agg[r[kn][:100]] += float(r[mv].replace(',', ''))
        cnt[r[kn][:100]] += 1
    except Exception:
        pass
tot = sum(agg.values())
print(f'total {tot / 1e3:.1f} us over {sum(cnt.values())} launches')
for k, v in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f'{v / tot * 100:6.2f}% {v / 1e3:10.1f}us x{cnt[k]:4d} {k}')
