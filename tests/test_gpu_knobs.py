"""GPU: configuration knobs of the hot path (tr_calling_config.*, rescaling.*) against the
oracle, and the per-read status / host-evaluation paths."""
import numpy as np
import pytest

from oracle import caller_oracle as co
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata
from warpstr_b200.config import CallerConfig, RescalerConfig

pytestmark = pytest.mark.gpu


def _run(cc, rc, name='HD', n=6, seed=70, noise=0.2, flank=110, engine=None):
    from warpstr_b200.caller import CallerEngine
    eng = CallerEngine(cc, rc)
    locus = synth.make_locus(name, seed=seed, flank_length=flank)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    ids = [eng.add_automaton(s, flank) for s in stas]
    reads = synth.make_reads(locus, n, seed=seed + 1, noise=noise)
    res = eng.call_batch([r.signal for r in reads], [ids[int(r.reverse)] for r in reads],
                         [r.reverse for r in reads], engine=engine)
    kn = co.Knobs(cc.min_values_per_state, cc.states_in_segment, rc.reps_as_one, rc.threshold, rc.max_std, rc.method)
    want = [co.run_read(r.signal, co.tables_from(stas[int(r.reverse)]), flank, r.reverse, kn, impl='c') for r in reads]
    return res, want


@pytest.mark.parametrize('mv', [2, 3, 4, 5, 6, 7, 8, 10])
@pytest.mark.parametrize('name', ['HD', 'DM2', 'CAN', 'RFC1'])
def test_min_values_per_state(built_lib, oracle_c, name, mv):
    """Any dwell > 1 on any automaton, as the reference accepts (config.py:115, caller.py:206-218,
    265-268): the whole two-pass call, bit-exact."""
    res, want = _run(CallerConfig(min_values_per_state=mv), RescalerConfig(), name=name, n=4)
    for g, w in zip(res, want):
        assert g.seq == w.seq and g.resc_seq == w.resc_seq
        assert g.cost == w.cost and g.resc_cost == w.resc_cost


@pytest.mark.parametrize('sis', [3, 6, 9])
def test_states_in_segment(built_lib, oracle_c, sis):
    res, want = _run(CallerConfig(states_in_segment=sis), RescalerConfig(), noise=0.3)
    for g, w in zip(res, want):
        assert g.resc_seq == w.resc_seq and g.resc_cost == w.resc_cost


@pytest.mark.parametrize('thr,mstd', [(0.3, 0.5), (0.5, 0.25), (0.8, 0.8)])
def test_rescaling_thresholds(built_lib, oracle_c, thr, mstd):
    res, want = _run(CallerConfig(), RescalerConfig(threshold=thr, max_std=mstd))
    for g, w in zip(res, want):
        assert g.resc_seq == w.resc_seq and g.resc_cost == w.resc_cost


def test_spline_with_interior_knots_goes_to_the_host(built_lib, oracle_c):
    """threshold > 1 admits pairs whose residual can exceed s = m: FITPACK then adds knots;
    the device flags the read (status 5) and the host evaluates it with scipy -- same result
    as the reference either way."""
    res, want = _run(CallerConfig(), RescalerConfig(threshold=3.0, max_std=3.0), noise=0.9, n=8)
    for g, w in zip(res, want):
        assert g.resc_seq == w.resc_seq
        assert g.resc_cost == pytest.approx(w.resc_cost, rel=1e-9)


@pytest.mark.parametrize('name', ['HD', 'DM2'])
def test_median_method_on_the_device(built_lib, oracle_c, name):
    """rescaling.method 'median': the run's state value is np.median of its samples."""
    res, want = _run(CallerConfig(), RescalerConfig(method='median'), name=name, n=8, noise=0.25)
    for g, w in zip(res, want):
        assert g.seq == w.seq and g.resc_seq == w.resc_seq
        assert g.cost == w.cost and g.resc_cost == w.resc_cost


@pytest.mark.parametrize('rc,engine', [(RescalerConfig(method='median'), 'host'), (RescalerConfig(reps_as_one=True), 'host')])
def test_host_engine_variants(built_lib, oracle_c, rc, engine):
    res, want = _run(CallerConfig(), rc, n=3, engine=engine)
    for g, w in zip(res, want):
        assert g.seq == w.seq and g.resc_seq == w.resc_seq
        assert g.resc_cost == pytest.approx(w.resc_cost, rel=1e-9)


@pytest.mark.parametrize('method', ['mean', 'median'])
@pytest.mark.parametrize('name', ['HD', 'DM2', 'C9ORF72_100'])
def test_reps_as_one_on_the_device(built_lib, oracle_c, name, method):
    """rescaling.reps_as_one: one alignment entry per distinct state over all its samples
    (caller.py:69-79) -- on the device, no host evaluation, bit-equal to the oracle."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('error', RuntimeWarning)          # the host path announces itself with one
        res, want = _run(CallerConfig(), RescalerConfig(reps_as_one=True, method=method), name=name, n=6, noise=0.25)
    for g, w in zip(res, want):
        assert g.seq == w.seq and g.resc_seq == w.resc_seq
        assert g.cost == w.cost and g.resc_cost == w.resc_cost


def test_invalid_mv_is_reported(built_lib):
    """The reference asserts min_values_per_state > 1 (config.py:115)."""
    from warpstr_b200 import _lib
    from warpstr_b200.caller import CallerEngine
    eng = CallerEngine(CallerConfig(min_values_per_state=4))
    eng.cc.min_values_per_state = 1
    locus = synth.make_locus('AAAT', seed=1)
    with pytest.raises(_lib.WarpstrError):
        eng.add_automaton(StateAutomata(locus.template_regex), 110)


def test_ttest_tie_guard_sends_reads_to_the_host_libm(built_lib, oracle_c):
    """d_ttest_ties: with the default width (16 ulp) no read of a batch is flagged; with an absurd width
    every read is, takes the host evaluation (the host's own pow decides, as in the reference) and still
    agrees with the oracle."""
    from warpstr_b200.caller import CallerEngine
    locus = synth.make_locus('HD', seed=70)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    reads = synth.make_reads(locus, 6, seed=71, noise=0.3)
    want = [co.run_read(r.signal, co.tables_from(stas[int(r.reverse)]), 110, r.reverse, impl='c') for r in reads]
    for width, flagged in ((0, 0), (1 << 51, len(reads))):
        eng = CallerEngine()
        eng.ttest_guard_ulps = width
        ids = [eng.add_automaton(s, 110) for s in stas]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            res = eng.call_batch([r.signal for r in reads], [ids[int(r.reverse)] for r in reads],
                                 [r.reverse for r in reads])
        assert eng.last_ttest_ties == flagged
        for g, w in zip(res, want):
            assert g.seq == w.seq and g.resc_seq == w.resc_seq
            assert g.cost == w.cost and g.resc_cost == w.resc_cost
