"""Instruction counts / stall samples of one kernel by opcode and by contiguous SASS region:
    python tools/ncu_regions.py file.ncu-rep <kernel-id> [region-size]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', ] + (([] if sys.argv[2] == '-' else ['--kernel-id', ':::' + sys.argv[2]]) if not sys.argv[2].isalpha() and '_' not in sys.argv[2] else ['--kernel-name', 'regex:' + sys.argv[2]]), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
h = rows[hi]
iS, iE, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
data = [(idx, r[iS].strip(), int(r[iE]), int(r[iSamp] or 0)) for idx, r in enumerate(rows[hi + 1:]) if len(r) > iE and r[iE].isdigit()]
tot_e = sum(d[2] for d in data); tot_s = sum(d[3] for d in data)
print('instructions executed', tot_e, 'samples', tot_s, 'static instructions', len(data))
ops = collections.Counter(); ops_s = collections.Counter()
for idx, src, e, s in data:
    op = src.split()[1] if src.startswith('@') else src.split()[0]
    op = op.split('.')[0]
    ops[op] += e; ops_s[op] += s
for op, e in ops.most_common(22):
    print(f'  {op:12s} {100*e/tot_e:5.1f}% exec   {100*ops_s[op]/max(tot_s,1):5.1f}% samples')
R = int(sys.argv[3]) if len(sys.argv) > 3 else 100
print('regions of', R, 'instructions: start idx, exec %, sample %, FFMA2 share')
for i in range(0, len(data), R):
    chunk = data[i:i+R]
    e = sum(d[2] for d in chunk); s = sum(d[3] for d in chunk)
    f = sum(d[2] for d in chunk if 'FFMA2' in d[1])
    if e > 0.005 * tot_e:
        print(f'  #{chunk[0][0]:5d}  {100*e/tot_e:5.1f}%  {100*s/max(tot_s,1):5.1f}%  ffma2 {100*f/max(e,1):4.0f}%')
