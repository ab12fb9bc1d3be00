"""End-to-end (host buffers in, host arrays out) time of the bench batch for a few pipeline settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
locus = synth.make_locus('HD', seed=1)
stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
sig, off, lengths, rev, truth = synth.make_read_batch(locus, n, seed=2000)
eng = CallerEngine()
ids = [eng.add_automaton(s, 110) for s in stas]
aut = np.where(rev > 0, ids[1], ids[0]).astype(np.int32)
host = torch.from_numpy(sig).pin_memory()
d_sig = host.cuda()


def timed(fn, reps=6):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


print('resident', round(timed(lambda: eng.call_packed(d_sig, off, lengths, aut, rev)), 2), 'ms')
for chunk, lanes in ((25000, 2), (50000, 2), (25000, 3), (12500, 2), (100000, 1), (35000, 2)):
    ms = timed(lambda: eng.call_arrays(host, off, lengths, aut, rev, chunk_reads=chunk, lanes=lanes))
    print(f'call_arrays chunk {chunk} lanes {lanes}: {ms:.2f} ms  {n / ms / 1e3:.3f} M reads/s')
