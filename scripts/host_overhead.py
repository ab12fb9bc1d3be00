"""Host-side cost of one large wstr_call_batch (WSTR_DEBUG_TIMING=1 prints the library's own sections)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warpstr_b200 import synth, _lib
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
locus = synth.make_locus('HD', seed=1)
stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
sig, off, lengths, rev, truth = synth.make_read_batch(locus, 20000, seed=3)
rep = n // 20000
sig_t = torch.from_numpy(sig[:-2]).cuda().repeat(rep)
sig_t = torch.cat((sig_t, torch.zeros(2, dtype=torch.float64, device='cuda')))
stride = len(sig) - 2
off = np.concatenate([off + k * stride for k in range(rep)])
lengths = np.tile(lengths, rep); rev = np.tile(rev, rep)
eng = CallerEngine()
ids = [eng.add_automaton(s, 110) for s in stas]
aut = np.where(rev > 0, ids[1], ids[0]).astype(np.int32)
for it in range(6):
    t0 = time.perf_counter()
    o = eng.call_packed(sig_t, off, lengths, aut, rev, want_seq=False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'iter {it}: enqueue {1e3*(t1-t0):.1f} ms, gpu wait {1e3*(t2-t1):.1f} ms', flush=True)
print('no sync between calls:')
t0 = time.perf_counter()
for it in range(4):
    ta = time.perf_counter()
    o = eng.call_packed(sig_t, off, lengths, aut, rev, want_seq=False)
    print(f'  enqueue {1e3*(time.perf_counter()-ta):.1f} ms', flush=True)
torch.cuda.synchronize()
print(f'total {1e3*(time.perf_counter()-t0):.1f} ms for 4 calls')
