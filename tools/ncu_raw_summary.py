"""one line per kernel of an `ncu -i R.ncu-rep --page raw --csv` export: duration, DRAM bytes, SM / DRAM throughput, occupancy, registers"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(n):
    return hdr.index(n) if n in hdr else None


C = {'name': col('Kernel Name'), 'dur': col('gpu__time_duration.sum'), 'rd': col('dram__bytes_read.sum'), 'wr': col('dram__bytes_write.sum'),
     'sm': col('sm__throughput.avg.pct_of_peak_sustained_elapsed'), 'dram': col('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
     'warps': col('sm__warps_active.avg.pct_of_peak_sustained_active'), 'regs': col('launch__registers_per_thread'),
     'grid': col('launch__grid_size'), 'inst': col('smsp__inst_executed.sum')}
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
print('duration us | dram read MB | dram write MB | sm throughput % | dram throughput % | warps active % | regs | grid | warp instructions')
for r in data:
    dur = float(r[C['dur']].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(units[C['dur']], 1e-3)
    mb = lambda k: float(r[C[k]].replace(',', '')) * SCALE.get(units[C[k]], 1e-6)
    print('%-52s %8.1f %9.2f %9.2f %6.1f %6.1f %6.1f %4s %6s %s' % (r[C['name']].split('(')[0][-52:], dur, mb('rd'), mb('wr'), float(r[C['sm']]),
          float(r[C['dram']]), float(r[C['warps']]), r[C['regs']], r[C['grid']], r[C['inst']].split('.')[0]))
