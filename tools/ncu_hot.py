"""Top warp-stall locations of an `ncu --page source --csv` dump (SASS view): python tools/ncu_hot.py dump.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
si, ss, ie = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ss]), k, r[si].strip(), float(r[ie]), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1.0
print('total samples', tot, 'instructions', len(data))
for s, k, src, n, r in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    why = sorted(((float(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f'{s / tot * 100:5.1f}%  #{k:5d} exec {n:10.0f}  {src[:70]:70s} {why[0][1]}:{why[0][0]:.0f} {why[1][1]}:{why[1][0]:.0f}')
