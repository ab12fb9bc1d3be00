"""Quick device-resident timing of the DP pass (development probe, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warpstr_b200 import synth, _lib
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine, pack_signals

name = sys.argv[1] if len(sys.argv) > 1 else 'HD'
n_unique = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
locus = synth.make_locus(name, seed=1)
stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
eng = CallerEngine()
ids = [eng.add_automaton(s, 110) for s in stas]
for a in eng.automata: print(a.info())
reads = synth.make_reads(locus, n_unique, seed=2)
sel = np.arange(n_reads) % n_unique
sigs = [reads[i].signal for i in sel]
aut = np.array([ids[int(reads[i].reverse)] for i in sel], dtype=np.int32)
host, off, lengths = pack_signals(sigs)
d_sig = host.cuda()
need = _lib.warp_workspace_bytes(eng.automata, aut, lengths)
print('reads', n_reads, 'samples', int(lengths.sum()), 'workspace GB', need / 1e9)
ws = torch.empty(min(need, 60 << 30), dtype=torch.uint8, device='cuda')
d_trace = torch.empty(d_sig.numel(), dtype=torch.int32, device='cuda')
d_status = torch.zeros(n_reads, dtype=torch.int32, device='cuda')
S = np.array([stas[0].n_states, stas[1].n_states])
cells = float((lengths.astype(np.int64) * S[aut - ids[0]]).sum())
for rep in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    _lib.warp_batch(eng.automata, aut, d_sig, off, lengths, None, None, ws, d_trace, None, d_status)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'rep {rep}: {ms:.2f} ms  {cells / ms / 1e6:.1f} GCUPS (fill+traceback)  {n_reads / ms * 1e3:.0f} reads/s/pass')
print('fp64 add rate T/s', _lib.measure_fp64_add_rate())
