"""GPU: the BASELINE configurations as parity cases at (near) full size, checked through
size-independent properties plus oracle spot checks.

  C2  HD batch                      -> lengths == simulated truth, wave-partition invariance, idempotence
  C3  FMR1 / (MGG) / DM2            -> oracle on a sample, truth on all
  C4  C9orf72 x1000 (T ~ 56-60k)    -> traceback memory stress, oracle on one read
  C5  multi-locus panel             -> many automata in one batch
"""
import numpy as np
import pytest

from oracle import caller_oracle as co
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from warpstr_b200.caller import CallerEngine
    return CallerEngine(**kw)


def _batch(eng, name, n, seed, noise=0.15):
    locus = synth.make_locus(name, seed=seed)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    ids = [eng.add_automaton(s, 110) for s in stas]
    sig, off, lengths, rev, truth = synth.make_read_batch(locus, n, seed=seed + 1, noise=noise)
    aut = np.where(rev > 0, ids[1], ids[0]).astype(np.int32)
    return locus, stas, (sig, off, lengths, rev, truth, aut)


def _call(eng, sig, off, lengths, rev, aut, **kw):
    import torch
    d = torch.from_numpy(sig).cuda()
    o = eng.call_packed(d, off, lengths, aut, rev, **kw)
    return {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in o.items()}


def test_c2_hd_batch_properties(built_lib, oracle_c):
    eng = _engine()
    locus, stas, (sig, off, lengths, rev, truth, aut) = _batch(eng, 'HD', 20000, seed=100)
    a = _call(eng, sig, off, lengths, rev, aut, want_debug=True)
    assert not a['status'].any()
    # the caller recovers the simulated allele length of (nearly) every read
    assert (a['len2'] == truth).mean() > 0.995
    # a trace is a walk over the automaton: starts in one of the row-0 states, ends in the end state
    t2 = a['trace2']
    firsts = t2[off]
    lasts = t2[off + lengths - 1]
    assert (firsts <= 4).all()
    ends = np.array([stas[0].endstate, stas[1].endstate])
    assert (lasts == ends[rev.astype(int)]).all()
    # idempotence and wave-partition invariance: a workspace that forces ~8 waves gives the same bits
    b = _call(eng, sig, off, lengths, rev, aut, want_debug=True)
    small = _engine(workspace_bytes=int(1.5e9))
    for s in stas:
        small.add_automaton(s, 110)
    c = _call(small, sig, off, lengths, rev, aut, want_debug=True)
    for k in ('len1', 'len2', 'cost1', 'cost2', 'trace1', 'trace2', 'rescaled', 'seq2'):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    # oracle spot check on reads spread over the batch
    tbs = [co.tables_from(s) for s in stas]
    for r in (0, 777, 4242, 19999):
        x = sig[off[r]:off[r] + lengths[r]]
        want = co.run_read(x, tbs[int(rev[r])], 110, bool(rev[r]), impl='c')
        assert np.array_equal(a['trace2'][off[r]:off[r] + lengths[r]], want.trace2)
        assert a['cost2'][r] == want.resc_cost and a['len2'][r] == len(want.resc_seq)


@pytest.mark.parametrize('name', ['FMR1', 'FMR1_MGG', 'DM2'])
def test_c3_interruptions_and_ambiguous_bases(built_lib, oracle_c, name):
    eng = _engine()
    locus, stas, (sig, off, lengths, rev, truth, aut) = _batch(eng, name, 3000, seed=200)
    a = _call(eng, sig, off, lengths, rev, aut)
    assert (a['status'] != 0).mean() < 0.01
    ok = a['status'] == 0
    assert (a['len2'][ok] == truth[ok]).mean() > 0.9
    tbs = [co.tables_from(s) for s in stas]
    for r in range(0, 3000, 500):
        if a['status'][r]:
            continue
        x = sig[off[r]:off[r] + lengths[r]]
        want = co.run_read(x, tbs[int(rev[r])], 110, bool(rev[r]), impl='c')
        assert a['len2'][r] == len(want.resc_seq) and a['len1'][r] == len(want.seq)
        assert a['cost1'][r] == want.cost and a['cost2'][r] == want.resc_cost
        s0 = int(a['seq_off'][r])
        assert a['seq2'][s0:s0 + a['len2'][r]].tobytes().decode() == want.resc_seq


def test_c4_long_expansion(built_lib, oracle_c):
    eng = _engine()
    locus, stas, (sig, off, lengths, rev, truth, aut) = _batch(eng, 'C9ORF72_1000', 48, seed=300)
    assert lengths.max() > 50000
    a = _call(eng, sig, off, lengths, rev, aut)
    assert not a['status'].any()
    assert (np.abs(a['len2'] - truth) <= 12).mean() > 0.9      # within two repeat units
    r = int(np.argmax(lengths))
    x = sig[off[r]:off[r] + lengths[r]]
    want = co.run_read(x, co.tables_from(stas[int(rev[r])]), 110, bool(rev[r]), impl='c')
    assert a['len2'][r] == len(want.resc_seq) and a['cost2'][r] == want.resc_cost


def test_c5_multi_locus_panel(built_lib, oracle_c):
    """Many loci (different kernel shapes) in one call."""
    eng = _engine()
    names = ['AAAT', 'HD', 'FMR1', 'DM2', 'CAN', 'RFC1', 'C9ORF72_100', 'FMR1_MGG']
    sigs, auts, revs, truths, tbs = [], [], [], [], []
    for n in range(24):
        name = names[n % len(names)]
        locus = synth.make_locus(name, seed=400 + n)
        stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
        ids = [eng.add_automaton(s, 110) for s in stas]
        for rd in synth.make_reads(locus, 40, seed=500 + n):
            sigs.append(rd.signal); auts.append(ids[int(rd.reverse)]); revs.append(rd.reverse)
            truths.append(rd.truth_len); tbs.append((stas[int(rd.reverse)], rd.reverse))
    res = eng.call_batch(sigs, auts, revs)
    got = np.array([len(r.resc_seq) for r in res])
    assert (got == np.array(truths)).mean() > 0.95
    for r in range(0, len(sigs), 97):
        sta, rv = tbs[r]
        want = co.run_read(sigs[r], co.tables_from(sta), 110, rv, impl='c')
        assert res[r].resc_seq == want.resc_seq and res[r].resc_cost == want.resc_cost
