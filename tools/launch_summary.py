"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: per-kernel totals of the last step."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))


def is_step_start(name):
    """A step starts at the first-layer block (conv_i8_kernel<.., 1, 1, 128, 1, ..>: fp32 frames are packed just before it, packed
    u8 frames arrive from the host already in that form)."""
    return 'conv_i8_kernel<' in name and ', 1, 1, 128, 1,' in name

hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
seq = []
for r in rows[hdr + 2:]:
    if len(r) > vi:
        try:
            seq.append((r[ki], float(r[vi].replace(',', '')), r[gi]))
        except ValueError:
            pass
idx = [i for i, s in enumerate(seq) if is_step_start(s[0])]
step = seq[idx[-2]:idx[-1]] if len(idx) > 1 else seq
agg = collections.OrderedDict()
for n, v, g in step:
    k = n[:78]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
print('step total %.3f ms, %d launches' % (sum(v for _, v, _ in step) / 1e6, len(step)))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{v / 1e6:9.3f} ms {c:4d}  {k}')
if len(sys.argv) > 3:
    for n, v, g in step:
        if any(w in n for w in sys.argv[3].split(',')):
            print(f'{v / 1e6:8.3f}', g, n[20:100])
