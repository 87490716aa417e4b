"""Hottest SASS instructions (by stall samples) of one kernel in an .ncu-rep captured with --import-source on:
    python tools/ncu_hot.py file.ncu-rep <kernel-id> [top]"""
import csv
import io
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', *(['--kernel-id', ':::' + sys.argv[2]] if sys.argv[2].isdigit() else ['--kernel-name', 'regex:' + sys.argv[2]])],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
h = rows[hi]
iS, iE, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_')]
data = []
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) <= iE or not r[iE].isdigit():
        continue
    data.append((int(r[iSamp] or 0), int(r[iE]), idx, r[iS].strip(), r))
ts = sum(d[0] for d in data)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print('total samples', ts)
for s, n, idx, src, r in sorted(data, key=lambda d: -d[0])[:top]:
    why = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols if r[i] not in ('', '0')), reverse=True)[:2]
    print(f'{100 * s / ts:5.1f}%  exec {n:9d}  #{idx:5d}  {src[:70]:70s} {why}')
