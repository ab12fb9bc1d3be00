"""Per-chunk timeline of CallerEngine.call_arrays (H2D / call / D2H events) over 30 steps; prints the step
times with the library's kernel-time categories and the timelines of steps slower than 135 ms.  This is the
script that showed cudaMemGetInfo and copy-engine-queued memsets stalling the pipeline."""
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
for _ in range(3):
    eng.call_arrays(host, off, lengths, aut, rev)
gc.collect(); gc.disable()
slow = []
times = []
_lib.profile_enable(True); _lib.profile_read()
for rep in range(30):
    eng.timeline = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    host_marks = []
    eng.call_arrays(host, off, lengths, aut, rev)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    pr = _lib.profile_read()
    times.append((round(dt, 1), round(pr['plan_upload']['ms'], 1), round(pr['dp_fill_traceback']['ms'], 1), round(pr['midstage']['ms'], 1)))
    if dt > 135:
        tl = eng.timeline; e0 = tl[0][1]
        slow.append((dt, [(l, round(e0.elapsed_time(e), 1)) for l, e in tl]))
print(times)
for dt, tl in slow[:3]:
    print('SLOW %.1f' % dt)
    for l, t in tl: print('   %-18s %8.1f' % (l, t))
