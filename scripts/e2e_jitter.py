"""Step-time jitter of four ways to run the bench batch (whole batch resident, chunked resident, copy then
chunks, the pipelined call_arrays), 20 steps each, synchronising after every step."""
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
d_sig = host.cuda()
bounds = [0, 3125, 9375, 21875, 46875, 71875, 96875, 100000]
chunks = list(zip(bounds[:-1], bounds[1:]))
def v_resident_chunks():
    for a, b in chunks:
        lo = int(off[a]); hi = int(off[b - 1] + ((int(lengths[b - 1]) + 1) & ~1) + 2)
        eng.call_packed(d_sig[lo:hi], off[a:b] - lo, lengths[a:b], aut[a:b], rev[a:b])
def v_resident_whole():
    eng.call_packed(d_sig, off, lengths, aut, rev)
def v_e2e():
    eng.call_arrays(host, off, lengths, aut, rev)
def v_e2e2():
    eng.call_arrays(host, off, lengths, aut, rev, lanes=2)
def v_e2e2s():
    eng.call_arrays(host, off, lengths, aut, rev, lanes=2, chunk_reads=25000)
def v_h2d_then_chunks():
    d_sig.copy_(host, non_blocking=True)
    v_resident_chunks()
gc.collect(); gc.disable()
def v1(): eng.call_arrays(host, off, lengths, aut, rev, lanes=1, chunk_reads=50000)
for name, fn in (('resident whole', v_resident_whole), ('resident chunks', v_resident_chunks), ('e2e 1 lane, chunks to 50000', v1), ('e2e', v_e2e), ('e2e 2 lanes', v_e2e2), ('e2e 2 lanes, chunks to 25000', v_e2e2s)):
    for _ in range(3): fn()
    ts = []
    for rep in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    print(name, ts)
