"""Per-launch table of one step from an ncu launch list with several metrics (--metrics a,b,c --csv): the launches between two
consecutive pack_events_kernel launches.   python tools/launch_table.py gpurun_out/launches.csv [step index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))


def is_step_start(name):
    """A step starts at the first-layer block (conv_i8_kernel<.., 1, 1, 128, 1, ..>: fp32 frames are packed just before it, packed
    u8 frames arrive from the host already in that form)."""
    return 'conv_i8_kernel<' in name and ', 1, 1, 128, 1,' in name

h = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[h]
ki, mi, vi, ii, gi = (hdr.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Grid Size'))
d = {}
for r in rows[h + 1:]:
    if len(r) > vi:
        d.setdefault(r[ii], {'k': r[ki], 'g': r[gi]})[r[mi]] = r[vi]
ids = sorted(d, key=int)
pe = [i for i in ids if is_step_start(d[i]['k'])]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seg = [i for i in ids if int(pe[k]) <= int(i) < int(pe[k + 1])]
tot = 0.0
short = lambda m: m.split('.')[0].replace('sm__pipe_tensor_cycles_active', 'tensor%').replace('smsp__issue_active', 'issue%')
for i in seg:
    e = d[i]
    t = float(e['gpu__time_duration.sum'].replace(',', '')) / 1e3
    tot += t
    extra = '  '.join('%s %5.1f' % (short(m), float(v.replace(',', ''))) for m, v in e.items() if m not in ('k', 'g', 'gpu__time_duration.sum'))
    name = e['k'].replace('void ', '').replace('ss::<unnamed>::', '').replace('unnamed>::', '')
    print('%8.1f us  %s  grid %-14s %s' % (t, extra, e['g'], name[:90]))
print('step total %.1f us, %d launches' % (tot, len(seg)))
