import os, sys, time, gc
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from warpstr_b200 import _lib, synth, caller as cmod
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine
locus = synth.make_locus('HD', seed=1)
stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
eng = CallerEngine()
ids = [eng.add_automaton(s, locus.flank_length) for s in stas]
sig, off, lengths, rev, truth = synth.make_read_batch(locus, 100000, seed=2000)
aut = np.where(rev > 0, ids[1], ids[0]).astype(np.int32)
d_sig = torch.from_numpy(sig).cuda()
bounds = [0, 3125, 9375, 21875, 46875, 71875, 96875, 100000]
chunks = list(zip(bounds[:-1], bounds[1:]))
T = {}
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
    setattr(mod, name, g)
wrap(_lib, 'call_workspace_bytes'); wrap(_lib, 'call_batch'); wrap(torch.cuda, 'mem_get_info')
def step():
    for a, b in chunks:
        lo = int(off[a]); hi = int(off[b - 1] + ((int(lengths[b - 1]) + 1) & ~1) + 2)
        t0 = time.perf_counter()
        eng.call_packed(d_sig[lo:hi], off[a:b] - lo, lengths[a:b], aut[a:b], rev[a:b])
        T.setdefault('call_packed', []).append((time.perf_counter() - t0) * 1e3)
gc.collect(); gc.disable()
for _ in range(3): step()
torch.cuda.synchronize(); T.clear()
ts = []
for rep in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); ts.append((round((time.perf_counter() - t0) * 1e3, 1), round((t1 - t0) * 1e3, 1)))
print('step (total, host enqueue)', ts)
for k, v in T.items():
    v = np.array(v); print('%-22s n=%d median=%.2f p99=%.2f max=%.2f  top5=%s' % (k, len(v), np.median(v), np.percentile(v, 99), v.max(), np.round(np.sort(v)[-5:], 1)))
