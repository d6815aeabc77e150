"""Table of the conv_i8_kernel launches of one forward from an `ncu --set full` capture exported with
    ncu -i rep.ncu-rep --page raw --csv > raw.csv
Prints time, DRAM bytes, tensor-pipe / issue activity, L2 and L1 throughput per launch and the DRAM bytes of the step; with
--json KEY FILE it also records the total under KEY in profiles/traffic.json (what bench.py reports as roofline.traffic)."""
import csv
import json
import sys

COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'lts__t_sector_hit_rate.pct']
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, units = rows[h], rows[h + 1]
ki = hdr.index('Kernel Name')
ci = [hdr.index(c) for c in COLS]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


print(' | '.join(['Kernel Name'] + COLS))
tot_bytes, tot_us = 0.0, 0.0
for r in rows[h + 2:]:
    if len(r) <= max(ci) or 'conv_i8_kernel' not in r[ki]:
        continue
    name = r[ki].replace('void ', '').replace('ss::<unnamed>::', '').replace('(ss::<unnamed>::I8Params)', '')
    vals = [r[i] for i in ci]
    rd, wr = to_bytes(r[ci[1]], units[ci[1]]), to_bytes(r[ci[2]], units[ci[2]])
    t = float(r[ci[0]].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(units[ci[0]], 1.0)
    tot_bytes += rd + wr
    tot_us += t
    print(' | '.join([name[:60], '%.1f us' % t, '%.2f MB' % (rd / 1e6), '%.2f MB' % (wr / 1e6)] + ['%.1f' % float(v.replace(',', '')) for v in vals[3:]]))
print('# %d launches, %.1f us, DRAM read + write %.1f MB' % (sum(1 for r in rows[h + 2:] if len(r) > ki and 'conv_i8_kernel' in r[ki]), tot_us, tot_bytes / 1e6))
if '--json' in sys.argv:
    key, path = sys.argv[sys.argv.index('--json') + 1], sys.argv[sys.argv.index('--json') + 2]
    tj = json.load(open(path))
    tj[key] = tot_bytes
    tj.setdefault('sources', {})[key] = sys.argv[1]
    json.dump(tj, open(path, 'w'), indent=1)
