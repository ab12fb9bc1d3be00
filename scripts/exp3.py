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
d = torch.empty_like(host, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'H2D {host.numel()*8/1e9:.2f} GB in {dt*1e3:.1f} ms = {host.numel()*8/dt/1e9:.1f} GB/s')
d_sig = d
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('resident want_seq=True  %.1f ms' % t(lambda: eng.call_packed(d_sig, off, lengths, aut, rev, want_seq=True)))
print('resident want_seq=False %.1f ms' % t(lambda: eng.call_packed(d_sig, off, lengths, aut, rev, want_seq=False)))
for ch in (100000, 50000, 25000, 12500, 6250):
    print('e2e chunk %6d  %.1f ms' % (ch, t(lambda: eng.call_arrays(host, off, lengths, aut, rev, chunk_reads=ch))))
# host-side cost of one call without GPU work visible: time the python+C planning by calling on a tiny GPU batch? use cProfile
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); eng.call_packed(d_sig, off, lengths, aut, rev, want_seq=True); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
