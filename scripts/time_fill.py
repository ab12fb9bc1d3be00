#!/usr/bin/env python
"""Time the DP fill(+traceback) kernel alone on the bench batch (first pass; optional masked second pass).
WSTR_LIB selects a library variant.  Prints one line: ms per launch, GCUPS, rows/s, cycles per row per SMSP."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from warpstr_b200 import _lib, synth  # noqa: E402
from warpstr_b200.automata import StateAutomata  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--reads', type=int, default=100000)
ap.add_argument('--locus', default='HD')
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--tag', default=os.environ.get('WSTR_LIB', 'stock'))
args = ap.parse_args()

locus = synth.make_locus(args.locus, seed=1)
stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
auts = [_lib.DeviceAutomaton.from_automaton(s, locus.flank_length, 4) for s in stas]
sig, off, lengths, rev, truth = synth.make_read_batch(locus, args.reads, seed=2000)
aut = rev.astype(np.int32)
d_sig = torch.from_numpy(sig).cuda()
need = _lib.warp_workspace_bytes(auts, aut, lengths)
ws = torch.empty(need, dtype=torch.uint8, device='cuda')
d_trace = torch.zeros(d_sig.numel(), dtype=torch.int32, device='cuda')
d_status = torch.zeros(args.reads, dtype=torch.int32, device='cuda')
n_states = np.array([s.n_states for s in stas])
cells = float((lengths.astype(np.int64) * n_states[aut]).sum())
rows = float(lengths.sum())
_lib.warp_batch(auts, aut, d_sig, off, lengths, None, None, ws, d_trace, None, d_status)
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_read()
for _ in range(args.reps):
    _lib.warp_batch(auts, aut, d_sig, off, lengths, None, None, ws, d_trace, None, d_status)
torch.cuda.synchronize()
p = _lib.profile_read()['dp_fill_traceback']
ms = p['ms'] / p['launches']
cyc = ms * 1e-3 * 1.965e9 * 148 * 4 / rows
print(f'{args.tag}: {ms:.2f} ms/launch  {cells / ms / 1e6:.0f} GCUPS  {cyc:.1f} cycles/row/SMSP  '
      f'checksum {int(d_trace.sum().item())} bad {int((d_status != 0).sum().item())}')
