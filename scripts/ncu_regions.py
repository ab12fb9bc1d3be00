#!/usr/bin/env python
"""Summarise an ncu report's source page: code regions (runs of SASS with similar execution counts),
their share of executed instructions and of stall samples, plus the raw-page key metrics."""
import csv, subprocess, sys, io
rep = sys.argv[1]
sel = []
if len(sys.argv) > 2 and sys.argv[2].startswith('#'):      # '#n': the n-th kernel of a multi-kernel report
    sel = ['--launch-skip', sys.argv[2][1:], '--launch-count', '1']
    del sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
print('kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(vals[i] or 0) > 0.05):
        print(f'{h:90s} {vals[i]:>16s} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith('stall_')]
ti = sum(int(r[ix['Instructions Executed']] or 0) for r in data); ts = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total inst', ti, 'samples', ts)
runs = []; cur = None
for k, r in enumerate(data):
    e = int(r[ix['Instructions Executed']] or 0); sm = int(r[ix['# Samples']] or 0)
    if cur and abs(e - cur['e0']) <= 0.25 * max(cur['e0'], 1):
        cur['n'] += 1; cur['inst'] += e; cur['samp'] += sm; cur['end'] = k
    else:
        cur = {'start': k, 'end': k, 'e0': e, 'n': 1, 'inst': e, 'samp': sm}; runs.append(cur)
    for h in stalls:
        cur[h] = cur.get(h, 0) + int(r[ix[h]] or 0)
for c in runs:
    if c['inst'] > 0.004 * ti or c['samp'] > 0.01 * ts:
        top = sorted(((c[h], h) for h in stalls), reverse=True)[:5]
        tops = ' '.join(f"{h[6:]}={100*v/max(c['samp'],1):.0f}%" for v, h in top)
        print(f"{data[c['start']][0][-5:]}..{data[c['end']][0][-5:]} n={c['n']:4d} exec={c['e0']:>10} inst%={100*c['inst']/ti:5.1f} samp%={100*c['samp']/ts:5.1f}  {tops}")
if len(sys.argv) > 2:   # dump a region instruction by instruction: start end (hex suffixes)
    a, b = sys.argv[2], sys.argv[3]
    on = False
    for r in data:
        if r[0].endswith(a): on = True
        if on:
            st = ' '.join(f"{h[6:]}={r[ix[h]]}" for h in stalls if int(r[ix[h]] or 0) > 0)
            print(r[0][-5:], f"{r[ix['# Samples']]:>6}", r[ix['Source']][:70].ljust(70), st)
        if r[0].endswith(b): break
