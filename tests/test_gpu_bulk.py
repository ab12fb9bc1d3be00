"""GPU: bulk parity (10^3..10^4 reads through the fast oracle).  Every read of a batch goes
through wstr_call_batch and through the oracle (DP in C, bulk mid-stage, a process pool on the host
cores); zero mismatches in lengths, sequences, costs and both traces are required.

WSTR_BULK_READS sets the reads per configuration (default 500: eight configurations, 4 000 reads, about
a minute on 16 cores; the round's full run used 2 000 = 16 000 reads)."""
import os

import numpy as np
import pytest

from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata

pytestmark = pytest.mark.gpu

N_READS = int(os.environ.get('WSTR_BULK_READS', '500'))


@pytest.mark.parametrize('noise', [0.15, 0.30])
@pytest.mark.parametrize('name', ['HD', 'FMR1', 'DM2', 'CAN'])
def test_bulk_parity(built_lib, oracle_c, name, noise):
    import torch
    from oracle import bulk
    from warpstr_b200.caller import CallerEngine
    locus = synth.make_locus(name, seed=900)
    regexes = [locus.template_regex, locus.reverse_regex]
    sig, off, lengths, rev, truth = synth.make_read_batch(locus, N_READS, seed=901, noise=noise)
    signals = [sig[o:o + n] for o, n in zip(off, lengths)]
    want = bulk.run_reads(regexes, 110, signals, rev.astype(int), rev.astype(bool), want_traces=True)

    eng = CallerEngine()
    ids = [eng.add_automaton(StateAutomata(rx), 110) for rx in regexes]
    aut = np.where(rev > 0, ids[1], ids[0]).astype(np.int32)
    o = eng.call_packed(torch.from_numpy(sig).cuda(), off, lengths, aut, rev, want_debug=True)
    g = {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in o.items()}

    bad = {'status': 0, 'len': 0, 'seq': 0, 'cost': 0, 'trace': 0}
    ties = int((g['ties'] > 0).sum())
    for r, w in enumerate(want):
        if w[0] == 'error':                        # the reference raises for this read: the device must flag it
            bad['status'] += int(g['status'][r] == 0)
            continue
        if g['status'][r] != 0:
            bad['status'] += 1
            continue
        if g['ties'][r] > 0:                       # decided by the host's libm (see d_ttest_ties): not compared here
            continue
        seq1, seq2, c1, c2, t1, t2 = w
        s0 = int(g['seq_off'][r])
        bad['len'] += int(g['len1'][r] != len(seq1) or g['len2'][r] != len(seq2))
        bad['seq'] += int(g['seq1'][s0:s0 + g['len1'][r]].tobytes().decode() != seq1 or
                          g['seq2'][s0:s0 + g['len2'][r]].tobytes().decode() != seq2)
        bad['cost'] += int(g['cost1'][r] != c1 or g['cost2'][r] != c2)
        a, n = int(off[r]), int(lengths[r])
        bad['trace'] += int(not np.array_equal(g['trace1'][a:a + n], t1) or not np.array_equal(g['trace2'][a:a + n], t2))
    print(f'bulk parity {name} noise {noise}: {N_READS} reads, mismatches {bad}, t-test ties {ties}')
    assert not any(bad.values()), bad
    assert ties <= max(1, N_READS // 1000)         # a near-tie is a ~1e-11 event per read
