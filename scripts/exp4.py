import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from warpstr_b200 import _lib, synth
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine
locus = synth.make_locus('HD', seed=1)
stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
eng = CallerEngine()
ids = [eng.add_automaton(s, locus.flank_length) for s in stas]
sig, off, lengths, rev, truth = synth.make_read_batch(locus, 100000, seed=2000)
aut = np.where(rev > 0, ids[1], ids[0]).astype(np.int32)
host = torch.from_numpy(sig).pin_memory()
for ch in (25000, 12500):
    for rep in range(3):
        eng.timeline = [] if rep == 2 else None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        host_t = []
        eng.call_arrays(host, off, lengths, aut, rev, chunk_reads=ch)
        torch.cuda.synchronize(); print('chunk', ch, 'rep', rep, '%.1f ms' % ((time.perf_counter() - t0) * 1e3))
    tl = eng.timeline
    e0 = tl[0][1]
    for label, e in tl:
        print('   %-18s %8.2f ms' % (label, e0.elapsed_time(e)))
# host time of a call on a 25000 chunk
d_sig = host[:int(off[25000])].cuda()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); eng.call_packed(d_sig, off[:25000], lengths[:25000], aut[:25000], rev[:25000]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print('call_packed(25000): host %.2f ms, total %.2f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
