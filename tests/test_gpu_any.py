"""GPU parity of the catch-all DP kernel (csrc/dtw_any.cu): every dwell setting
(min_values_per_state > 1, reference src/config.py:115) on every automaton shape, bit-exact
against the C oracle -- and bit-exact against the specialised register kernels where both exist."""
import numpy as np
import pytest

from oracle import caller_oracle as co
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata
from warpstr_b200.config import CallerConfig

pytestmark = pytest.mark.gpu


def _engine(mv, generic):
    from warpstr_b200 import _lib
    from warpstr_b200.caller import CallerEngine
    eng = CallerEngine(CallerConfig(min_values_per_state=mv))
    eng._generic = generic
    return eng


def _add(eng, sta, flank):
    from warpstr_b200 import _lib
    _lib.set_generic_only(bool(eng._generic))
    try:
        return eng.add_automaton(sta, flank)
    finally:
        _lib.set_generic_only(False)


def _masks(reads, seed):
    rng = np.random.default_rng(seed)
    out = []
    for r in reads:
        m = np.zeros(len(r.signal), dtype=bool)
        for _ in range(6):
            a = int(rng.integers(200, max(201, len(m) - 200)))
            m[a:a + int(rng.integers(20, 120))] = True
        m[rng.integers(0, len(m), 30)] = True
        out.append(m)
    return out


@pytest.mark.parametrize('mv', [2, 3, 4, 5, 6, 7, 8, 11])
@pytest.mark.parametrize('name', ['HD', 'DM2', 'CAN', 'RFC1'])
def test_catch_all_traces_match_oracle(built_lib, oracle_c, name, mv):
    eng = _engine(mv, generic=True)
    locus = synth.make_locus(name, seed=41)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    ids = [_add(eng, s, 110) for s in stas]
    assert all(eng.automata[i].info()['states_per_lane'] == 0 for i in ids)      # really the catch-all
    reads = synth.make_reads(locus, 4, seed=43, noise=0.2)
    sigs = [r.signal for r in reads]
    aut = [ids[int(r.reverse)] for r in reads]
    traces, costs = eng.warp_batch(sigs, aut, return_end_cost=True)
    masks = _masks(reads, 5)
    traces_m = eng.warp_batch(sigs, aut, masks)
    for r, t, c, tm, m in zip(reads, traces, costs, traces_m, masks):
        tb = co.tables_from(stas[int(r.reverse)])
        m0 = np.zeros(len(r.signal), dtype=bool)
        assert np.array_equal(t, oracle_c.warp(r.signal, tb, m0, mv, 110)), (name, mv, r.name)
        assert c == oracle_c.fill(r.signal, tb, m0, mv, 110)[-1, tb.endstate]
        assert np.array_equal(tm, oracle_c.warp(r.signal, tb, m, mv, 110)), (name, mv, r.name, 'masked')


@pytest.mark.parametrize('mv', [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize('name', ['HD', 'FMR1_MGG', 'DM2', 'CAN', 'RFC1', 'C9ORF72_100'])
def test_catch_all_equals_specialised(built_lib, name, mv):
    """The same reads through both kernels: identical traces and end costs, first and masked pass."""
    locus = synth.make_locus(name, seed=47)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    reads = synth.make_reads(locus, 6, seed=49, noise=0.25)
    sigs = [r.signal for r in reads]
    masks = _masks(reads, 7)
    got = []
    for generic in (False, True):
        eng = _engine(mv, generic)
        ids = [_add(eng, s, 110) for s in stas]
        aut = [ids[int(r.reverse)] for r in reads]
        t, c = eng.warp_batch(sigs, aut, return_end_cost=True)
        got.append((t, c, eng.warp_batch(sigs, aut, masks)))
    for a, b in zip(got[0][0], got[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(got[0][1], got[1][1])
    for a, b in zip(got[0][2], got[1][2]):
        assert np.array_equal(a, b)


def test_catch_all_takes_what_no_layout_holds(built_lib, oracle_c):
    """In-degree 5 (an optional IUPAC base inside a loop: 519 states) and a long plain insert between
    two repeats (535 states): beyond the 512 register-resident positions / 4 candidates per state."""
    eng = _engine(4, generic=False)
    rng = np.random.default_rng(3)
    left, right = synth.random_flank(rng, 110), synth.random_flank(rng, 110)
    insert = synth.random_flank(rng, 300)
    cases = (('(A{N})', (('A', 20, 40, 'AC', 0.3),), 5),
             ('(CAG)' + insert + '(CTG)', (('CAG', 8, 12), insert, ('CTG', 8, 12)), 2))
    for regex, units, indeg in cases:
        sta = StateAutomata(left + regex + right)
        assert sta.n_states > 512 and int(np.diff(sta.in_ptr).max()) == indeg
        aid = eng.add_automaton(sta, 110)
        assert eng.automata[aid].info()['states_per_lane'] == 0
        locus = synth.SynthLocus('x', regex, left, right, units)
        reads = [r for r in synth.make_reads(locus, 8, seed=5, noise=0.2) if not r.reverse][:3]
        assert reads
        sigs = [r.signal for r in reads]
        traces = eng.warp_batch(sigs, [aid] * len(reads))
        masks = _masks(reads, 11)
        traces_m = eng.warp_batch(sigs, [aid] * len(reads), masks)
        tb = co.tables_from(sta)
        for r, t, tm, m in zip(reads, traces, traces_m, masks):
            assert np.array_equal(t, oracle_c.warp(r.signal, tb, np.zeros(len(r.signal), dtype=bool), 4, 110))
            assert np.array_equal(tm, oracle_c.warp(r.signal, tb, m, 4, 110))


def test_catch_all_in_one_batch_with_specialised_and_in_waves(built_lib, oracle_c):
    from warpstr_b200.caller import CallerEngine
    eng = CallerEngine(workspace_bytes=8 << 20)
    eng._generic = False
    sigs, aut, want = [], [], []
    for name, generic in (('HD', False), ('DM2', True), ('AAAT', True)):
        locus = synth.make_locus(name, seed=33)
        stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
        eng._generic = generic
        ids = [_add(eng, s, 110) for s in stas]
        for r in synth.make_reads(locus, 5, seed=133):
            sigs.append(r.signal); aut.append(ids[int(r.reverse)])
            want.append(oracle_c.warp(r.signal, co.tables_from(stas[int(r.reverse)]),
                                      np.zeros(len(r.signal), dtype=bool), 4, 110))
    for t, w in zip(eng.warp_batch(sigs, aut), want):
        assert np.array_equal(t, w)
    # too short for the dwell: same status as the specialised kernels
    with pytest.raises(IndexError):
        eng.warp_batch([sigs[-1][:4]], [aut[-1]])
