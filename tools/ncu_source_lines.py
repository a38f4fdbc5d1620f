"""Rank the CUDA source lines of one kernel by stall samples / executed instructions from
`ncu -i R.ncu-rep --page source --csv --print-source sass,cuda --kernel-name regex:<k> --launch-count 1 > X.csv`
usage: python tools/ncu_source_lines.py X.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index('# Samples'), hdr.index('Instructions Executed')
lines = []
for r in rows:
    if r and r[0] not in ('', 'Line No'):
        try:
            lines.append((int(r[0]), r[1], int(r[iS] or 0), int(r[iI] or 0)))
        except ValueError:
            pass
ts, ti = sum(x[2] for x in lines) or 1, sum(x[3] for x in lines) or 1
print('kernel:', rows[0][1][:100])
print('source lines %d, stall samples %d, warp instructions %d' % (len(lines), ts, ti))
for x in sorted(lines, key=lambda x: -x[2])[:top]:
    print('%5d  samples %5.1f%%  instr %5.1f%%  %s' % (x[0], 100.0 * x[2] / ts, 100.0 * x[3] / ti, x[1].strip()[:130]))
