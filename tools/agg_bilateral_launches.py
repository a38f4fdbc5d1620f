"""per-kernel totals of one bilateral-filter call per workload from an `ncu --metrics gpu__time_duration.sum,dram__bytes_*` launch list
of `tools/bilateral_bench.py --profile` (3 calls per workload; the last call of each is summarised)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
iN, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows:
    per.setdefault(r[iID], {'name': r[iN]})[r[iM]] = float(r[iV].replace(',', ''))
L = list(per.values())
starts = [i for i, l in enumerate(L) if 'pl_embed' in l['name']]
for wl in range(len(starts) // 3):
    s = starts[wl * 3 + 2]
    e = starts[wl * 3 + 3] if wl * 3 + 3 < len(starts) else len(L)
    tot = 0
    print('--- workload', wl)
    agg = collections.OrderedDict()
    for l in L[s:e]:
        n = l['name'].split('(')[0]
        if n.startswith('void at::'):
            continue
        a = agg.setdefault(n, [0, 0, 0, 0])
        a[0] += 1
        a[1] += l['gpu__time_duration.sum'] / 1e3
        a[2] += l['dram__bytes_read.sum'] / 1e6
        a[3] += l['dram__bytes_write.sum'] / 1e6
    for n, a in agg.items():
        print('%-28s x%d %9.1f us  dram rd %8.2f MB wr %8.2f MB' % (n, a[0], a[1], a[2], a[3]))
        tot += a[1]
    print('total %.1f us (serialised, cold-cache: shares, not bench values)' % tot)
