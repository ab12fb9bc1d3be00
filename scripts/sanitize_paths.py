"""Small invocations of every kernel path, meant to be run under compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python scripts/sanitize_paths.py
Covers: specialised and catch-all DP kernels (first and masked pass), the mid-stage with the pipelined fit and the
tiled sort (one thousand-repeat read), reps_as_one, median, the three ingestion forms, a batch cut into slices and
waves, the split-halves resident call, the normalisation kernel's window path, fallback and median filter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warpstr_b200 import _lib, synth
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine, pack_signals
from warpstr_b200.config import CallerConfig, RescalerConfig


def batch(eng, name, n, seed, generic=False, noise=0.2):
    locus = synth.make_locus(name, seed=seed)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    _lib.set_generic_only(generic)
    ids = [eng.add_automaton(s, 110) for s in stas]
    _lib.set_generic_only(False)
    reads = synth.make_reads(locus, n, seed=seed + 1, noise=noise)
    return reads, [ids[int(r.reverse)] for r in reads]


eng = CallerEngine()
for name, generic in (('HD', False), ('DM2', False), ('CAN', False), ('HD', True)):
    reads, aut = batch(eng, name, 3, 5, generic)
    res = eng.call_batch([r.signal for r in reads], aut, [r.reverse for r in reads])
    print(name, generic, [len(x.resc_seq) for x in res])
reads, aut = batch(eng, 'C9ORF72_1000', 1, 7)
print('long read', len(reads[0].signal), len(eng.call_batch([reads[0].signal], aut, [reads[0].reverse])[0].resc_seq))
for rc in (RescalerConfig(reps_as_one=True), RescalerConfig(method='median'), RescalerConfig(reps_as_one=True, method='median')):
    e2 = CallerEngine(CallerConfig(min_values_per_state=3), rc)
    reads, aut = batch(e2, 'HD', 3, 9)
    print(rc, [len(x.resc_seq) for x in e2.call_batch([r.signal for r in reads], aut, [r.reverse for r in reads])])
# slices + waves + split halves
e3 = CallerEngine(workspace_bytes=24 << 20)
e3.split_small = (4, 100000)
reads, aut = batch(e3, 'HD', 24, 11)
print('small workspace', [len(x.resc_seq) for x in e3.call_batch([r.signal for r in reads], aut, [r.reverse for r in reads])][:6])
# ingestion forms
rng = np.random.default_rng(1)
raws, wins = [], []
for r in reads[:6]:
    raw, lo, hi = synth.to_raw_int16(rng, r.signal, pad=4096, spike_rate=2e-3)
    raws.append(raw); wins.append((lo, hi))
out = eng.call_raw_batch(raws, wins, aut[:6] if False else [eng.automata and a for a in batch(eng, 'HD', 6, 11)[1]], [r.reverse for r in reads[:6]])
print('raw chain', [len(x.resc_seq) for x in out])
# normalisation on its own: the window path, its fallback (bimodal read), a read cut off by the window, the median filter
from warpstr_b200.normalize import normalize_windows
r1 = (470 + 28 * rng.standard_normal(9000)).astype(np.int16)
r1[[5, 6, 4000, 4001, 8999]] = 1500
r2 = np.where(rng.random(7000) < 0.5, 260, 940).astype(np.int16)
for mode in ('Brute', 'None', 'median3'):
    o = normalize_windows([r1, r2, r1[:37]], [(100, 3000), (0, 99999), (0, 36)], mode)
    print('normalise', mode, [len(x) for x in o])
torch.cuda.synchronize()
print('done')
