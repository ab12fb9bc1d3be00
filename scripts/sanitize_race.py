"""Short reads through the kernels that synchronise through shared memory (specialised and catch-all DP, tiled
pair sort via a lowered shared-memory limit is not reachable here; see sanitize_paths.py for memcheck):
  compute-sanitizer --tool racecheck python scripts/sanitize_race.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warpstr_b200 import _lib, synth
from warpstr_b200.automata import StateAutomata
from warpstr_b200.caller import CallerEngine
eng = CallerEngine()
for name, generic in (('AAAT', False), ('AAAT', True), ('DM2', False)):
    locus = synth.make_locus(name, seed=5, flank_length=30)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    _lib.set_generic_only(generic)
    ids = [eng.add_automaton(s, 30) for s in stas]
    _lib.set_generic_only(False)
    reads = synth.make_reads(locus, 2, seed=6)
    res = eng.call_batch([r.signal for r in reads], [ids[int(r.reverse)] for r in reads], [r.reverse for r in reads])
    print(name, generic, [len(x.resc_seq) for x in res], [len(r.signal) for r in reads])
torch.cuda.synchronize()
print('done')
