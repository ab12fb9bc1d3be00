"""Build container only (needs /root/reference): the oracle and the product's host code
against the UNMODIFIED reference functions on fresh inputs.  Skipped on the GPU box, where
the committed goldens (made by oracle/make_golden.py from the same functions) stand in."""
import numpy as np
import pytest

from oracle import refshim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not refshim.available(), reason='reference tree not present')]


@pytest.fixture(scope='module')
def ref():
    return refshim.load()


def test_automaton_builder_against_reference(ref):
    from warpstr_b200.automata import StateAutomata
    from warpstr_b200 import synth
    rng = np.random.default_rng(11)
    pats = ['(AAAT)', '((CGG){AGG})', '(CAN)', '(AC{GT}(TA))', 'AC{G}{T}CA(GA)', '(RY)N(A)']
    for pat in pats:
        seq = synth.random_flank(rng, 40) + pat + synth.random_flank(rng, 40)
        a, b = ref.StateAutomata(seq), StateAutomata(seq)
        assert [s.kmer for s in a.states] == b.kmers
        assert [s.value for s in a.states] == list(b.values)
        assert [[p.idx for p in s.incoming] for s in a.states] == [list(b.incoming_of(i)) for i in range(b.n_states)]
        assert a.mask == b.mask and a.endstate == b.endstate


def test_oracle_against_reference_run(ref):
    from oracle import caller_oracle as co
    from warpstr_b200 import synth
    locus = synth.make_locus('FMR1', seed=77, flank_length=40)
    rd = synth.make_reads(locus, 1, seed=78)[0]
    sta = ref.StateAutomata(locus.reverse_regex if rd.reverse else locus.template_regex)
    w = ref.WarpSTR(40, sta.states, sta.endstate, sta.mask, None, rd.reverse, rd.name)
    want = w.run(rd.signal)
    got = co.run_read(rd.signal, co.tables_from(sta), 40, rd.reverse, impl='c')
    assert (got.seq, got.resc_seq, got.cost, got.resc_cost) == (want.seq, want.resc_seq, want.cost, want.resc_cost)
    m0 = np.full(len(rd.signal), False)
    D = w._calc_dtw_astates(rd.signal, sta.states, m0)
    assert np.array_equal(D, co.fill_scalar(rd.signal, co.tables_from(sta), m0, 4, 40))


def test_wrapper_helpers_against_reference(ref):
    from warpstr_b200.wrapper import CallerWrapper
    mine = CallerWrapper.__new__(CallerWrapper)
    theirs = ref.wrapper.CallerWrapper.__new__(ref.wrapper.CallerWrapper)
    for pat in ('((CAGG){CAGM})(CAGA)(CA)', '(AGC)AACAGCCGCCAC(CGC)', '(MGG)', 'AC(A{CN}T)G(CA)'):
        assert mine.break_into_units(pat) == theirs.break_into_units(pat)
        assert mine.reverse_uniq_sequence(pat) == theirs.reverse_uniq_sequence(pat)
    mine.units, mine.repeat_units, mine.offsets = mine.break_into_units('((CAGG){CAGM})(CAGA)(CA)')
    theirs.units, theirs.repeat_units, theirs.offsets = mine.units, mine.repeat_units, mine.offsets
    for seq in ('CAGGCAGGCAGACAGACA', 'CAGGCAGGCAGCCAGACACACA', 'TTT', ''):
        assert mine.collapse_repeats(seq) == theirs.collapse_repeats(seq)


def test_normalize_oracle_against_reference(ref):
    from oracle import normalize_oracle as no
    rng = np.random.default_rng(5)
    raw = rng.integers(200, 1100, size=5001).astype(np.int16)
    assert np.array_equal(no.brute_remove(raw), ref.Fast5.brute_remove(raw))
    assert np.array_equal(no.normalize_signal_mad(raw), ref.normalize_signal_mad(raw))
