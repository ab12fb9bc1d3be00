"""Timeline of one call_arrays (float64 windows) step: when each chunk's copy, call and result copy begin/end."""
import os, sys, time, gc
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
for _ in range(4):
    eng.call_arrays(host, off, lengths, aut, rev)
torch.cuda.synchronize()
eng.timeline = []
t0 = time.perf_counter()
eng.call_arrays(host, off, lengths, aut, rev)
torch.cuda.synchronize()
print('wall %.1f ms' % (1e3 * (time.perf_counter() - t0)))
tl = eng.timeline; e0 = tl[0][1]
for l, e in sorted(((l, e0.elapsed_time(e)) for l, e in tl), key=lambda t: t[1]):
    print('   %-18s %8.2f' % (l, e))
