#!/usr/bin/env python
"""Static look at the fill kernel's hot loop: finds the largest backward-branch loop bodies in the SASS of one
kernel instantiation and prints their instruction mix.  Usage: sass_loop.py [obj] [mangled-substring]"""
import collections, re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else 'build/dtw.cu.o'
pat = sys.argv[2] if len(sys.argv) > 2 else 'ILi6ELi2ELi2ELi4EE'
out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
blocks = out.split('Function : ')
for blk in blocks[1:]:
    name = blk.split('\n', 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in blk.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA(?:\.U)?(?:\.ANY)?\s+(?:[!U]*P\d,\s*)?0x([0-9a-f]+)', t)
        if m and 'BRA' in t:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_idx:
                loops.append((i - addr_idx[tgt] + 1, addr_idx[tgt], i))
    loops = sorted(l for l in loops if int(sys.argv[4] if len(sys.argv) > 4 else 250) <= l[0] <= 900)
    print(name, 'total', len(ins))
    for n, s, e in loops[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
        c = collections.Counter()
        for a, t in ins[s:e + 1]:
            op = t.split()[1] if t.startswith('@') else t.split()[0]
            c[op.split('.')[0]] += 1
        fp64 = c['DADD'] + c['DSETP']
        print(f'  loop {ins[s][0]:#x}..{ins[e][0]:#x}: {n} instr, FP64 {fp64}, other {n - fp64}:',
              ' '.join(f'{k}={v}' for k, v in c.most_common()))
