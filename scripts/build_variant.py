#!/usr/bin/env python
"""Build an experimental variant of the library: extra nvcc flags -> warpstr_b200/libwarpstr_b200.<name>.so
(select it at run time with WSTR_LIB=<path>).  Usage: build_variant.py NAME [nvcc flags ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
objdir = os.path.join(ROOT, 'build', 'variant_' + name)
os.makedirs(objdir, exist_ok=True)
objs, procs = [], []
only_aux = any('WSTR_NORM' in e for e in extra)
for unit, unit_flags, objname in g.units():
    src = os.path.join(g.CSRC, unit)
    obj = os.path.join(objdir, objname)
    objs.append(obj)
    # only dtw.cu (or aux.cu for the WSTR_NORM_* knobs) depends on the experiment knobs; reuse the stock objects for the rest
    stock = os.path.join(ROOT, 'build', objname)
    if (unit != ('aux.cu' if only_aux else 'dtw.cu')) and os.path.exists(stock) and not any('WARPS_PER_CTA' in e for e in extra):
        objs[-1] = stock
        continue
    procs.append(subprocess.Popen([g._nvcc()] + g.NVCC_FLAGS + unit_flags + extra + ['-c', src, '-o', obj]))
for p in procs:
    if p.wait() != 0:
        sys.exit(1)
out = os.path.join(ROOT, 'warpstr_b200', f'libwarpstr_b200.{name}.so')
subprocess.check_call([g._nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + objs + ['-lcudart'])
print(out)
