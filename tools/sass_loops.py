"""Static look at the loops of one kernel in the built library (no GPU needed): for every backward branch whose body holds
FFMA2s, the instruction mix of the body.   python tools/sass_loops.py <kernel-substring> [lib.so] [min-ffma2]"""
import re, subprocess, sys, collections
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else 'pixlzr-rust_b200/libpixlzr_b200.so'
minf = int(sys.argv[3]) if len(sys.argv) > 3 else 6
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print('==', name[:100], len(ins), 'instructions')
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                loops.append((addr_idx[tgt], i))
    for (s, e) in loops:
        body = ins[s:e + 1]
        nf = sum('FFMA2' in t for _, t in body)
        if nf < minf:
            continue
        # innermost only
        if any(s2 >= s and e2 <= e and (s2, e2) != (s, e) and sum('FFMA2' in t for _, t in ins[s2:e2 + 1]) >= minf for (s2, e2) in loops):
            continue
        ops = collections.Counter()
        for _, t in body:
            op = t.split()[1] if t.startswith('@') else t.split()[0]
            ops['MOV' if ('IMAD.MOV' in t or op == 'MOV') else op.split('.')[0]] += 1
        print(f'  loop {ins[s][0]:#x}..{ins[e][0]:#x}: {len(body)} instr, ' + ', '.join(f'{k} {v}' for k, v in ops.most_common(9)))
