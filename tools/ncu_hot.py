"""Top stall-sample SASS lines of one kernel from an ncu report (source page):  python tools/ncu_hot.py rep.ncu-rep kernel_regex [N]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stalls = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try:
        v = float(r[isamp])
    except (ValueError, IndexError):
        continue
    top = sorted(((float(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
    data.append((v, idx, r[isrc][:90], r[iex], ' '.join(f'{n}:{int(c)}' for c, n in top if c)))
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
for v, idx, sx, ex, st in sorted(data, key=lambda x: -x[0])[:topn]:
    print('%6.0f %5.1f%% #%-5d x%-9s %-90s %s' % (v, 100 * v / tot, idx, ex, sx, st))
