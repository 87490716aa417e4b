"""Instruction mix (by SASS opcode) and stall-sample share of one kernel in an .ncu-rep captured with --import-source on:
    python tools/ncu_opmix.py file.ncu-rep <kernel-id>"""
import collections
import csv
import io
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--kernel-id', ':::' + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
h = rows[hi]
iS, iE, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    t = r[iS].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith('@') else t[0]).rstrip(';')
    base = op.split('.')[0]
    if base in ('LDS', 'STS', 'LDG', 'STG'):
        base = '.'.join(op.split('.')[:1] + [x for x in op.split('.')[1:] if x.isdigit()])
    n = int(r[iE])
    ops[base] += n
    tot += n
    samp[base] += int(r[iSamp] or 0)
ts = max(1, sum(samp.values()))
print(rows[0][1][:100] if rows and len(rows[0]) > 1 else '')
print('warp-instructions executed:', tot)
for op, n in ops.most_common(28):
    print(f'{op:12s} {n:12d} {100 * n / tot:5.1f}%   stall samples {100 * samp[op] / ts:5.1f}%')
