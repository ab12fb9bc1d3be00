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


def test_bundled_reads_against_reference_run(ref):
    """Two of the reference's own bundled test reads (real nanopore signal, golden/c1_bundled.npz):
    the unmodified reference's normalisation and WarpSTR.run against what the fixture holds from
    the oracle -- which is what the GPU path is compared with in tests/test_gpu_c1.py."""
    import os
    from warpstr_b200 import templates as tmpl
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'c1_bundled.npz'))
    left, right, seq, F = str(z['left']), str(z['right']), str(z['sequence']), int(z['flank_length'])
    regex = {False: left + seq + right,
             True: tmpl.reverse_complement(right) + tmpl.reverse_uniq_sequence(seq) + tmpl.reverse_complement(left)}
    order = np.argsort(z['r_end_raw'] - z['l_start_raw'])[:2]            # the two shortest windows
    for i in order:
        raw = np.cumsum(z[f'raw_delta{i}'].astype(np.int64)).astype(np.int16)
        norm = ref.normalize_signal_mad(ref.Fast5.brute_remove(raw))       # Fast5.get_data_processed, fast5.py:45-57
        x = np.ascontiguousarray(norm[int(z['l_start_raw'][i]):int(z['r_end_raw'][i]) + 1])
        rev = bool(z['reverse'][i])
        sta = ref.StateAutomata(regex[rev])
        got = ref.WarpSTR(F, sta.states, sta.endstate, sta.mask, None, rev, str(z['names'][i])).run(x)
        assert got.seq == str(z['seq'][i]) and got.resc_seq == str(z['resc_seq'][i])
        assert got.cost == z['cost1'][i] and got.resc_cost == z['cost2'][i]
        assert len(got.resc_seq) in (40, 44)
