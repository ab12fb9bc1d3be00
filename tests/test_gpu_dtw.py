"""GPU parity of the DP fill + traceback (wstr_warp_batch) against the oracle: traces must be
identical sample by sample (integer output, bit-exact bar), end costs equal to the last bit
(same float64 additions in the same order)."""
import numpy as np
import pytest

from oracle import caller_oracle as co
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata

pytestmark = pytest.mark.gpu

LOCI = ['AAAT', 'HD', 'FMR1', 'FMR1_MGG', 'DM2', 'CAN', 'RFC1', 'C9ORF72_100']


@pytest.fixture(scope='module')
def engine(built_lib):
    from warpstr_b200.caller import CallerEngine
    return CallerEngine()


def _setup(engine, name, n, seed, flank=110, noise=0.15):
    locus = synth.make_locus(name, seed=seed, flank_length=flank)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    ids = [engine.add_automaton(s, flank) for s in stas]
    reads = synth.make_reads(locus, n, seed=seed + 100, noise=noise)
    return locus, stas, ids, reads


@pytest.mark.parametrize('name', LOCI)
def test_first_pass_traces_match_oracle(engine, oracle_c, name):
    locus, stas, ids, reads = _setup(engine, name, 6, seed=5)
    sigs = [r.signal for r in reads]
    aut = [ids[int(r.reverse)] for r in reads]
    traces, costs = engine.warp_batch(sigs, aut, return_end_cost=True)
    for r, t, c in zip(reads, traces, costs):
        tb = co.tables_from(stas[int(r.reverse)])
        m0 = np.zeros(len(r.signal), dtype=bool)
        want = oracle_c.warp(r.signal, tb, m0, 4, 110)
        assert np.array_equal(t, want), (name, r.name)
        D = oracle_c.fill(r.signal, tb, m0, 4, 110)
        assert c == D[-1, tb.endstate]


@pytest.mark.parametrize('name', ['AAAT', 'HD', 'DM2', 'CAN'])
def test_masked_pass_matches_oracle(engine, oracle_c, name):
    """Second-pass semantics: rows with mask=True allow dwell mv-1 (caller.py:218,265-268)."""
    locus, stas, ids, reads = _setup(engine, name, 4, seed=9, noise=0.25)
    rng = np.random.default_rng(3)
    sigs, aut, masks = [], [], []
    for r in reads:
        m = np.zeros(len(r.signal), dtype=bool)
        for _ in range(6):                      # a few masked stretches inside the repeat
            a = int(rng.integers(900, len(m) - 900))
            m[a:a + int(rng.integers(20, 120))] = True
        m[rng.integers(0, len(m), 30)] = True   # and isolated rows
        sigs.append(r.signal); aut.append(ids[int(r.reverse)]); masks.append(m)
    traces = engine.warp_batch(sigs, aut, masks)
    for r, t, m in zip(reads, traces, masks):
        tb = co.tables_from(stas[int(r.reverse)])
        want = oracle_c.warp(r.signal, tb, m, 4, 110)
        assert np.array_equal(t, want), (name, r.name)


def test_short_flank_and_small_automaton(engine, oracle_c):
    locus, stas, ids, reads = _setup(engine, 'AAAT', 5, seed=21, flank=30)
    sigs = [r.signal for r in reads]
    aut = [ids[int(r.reverse)] for r in reads]
    traces = engine.warp_batch(sigs, aut)
    for r, t in zip(reads, traces):
        tb = co.tables_from(stas[int(r.reverse)])
        want = oracle_c.warp(r.signal, tb, np.zeros(len(r.signal), dtype=bool), 4, 30)
        assert np.array_equal(t, want)


def test_mixed_loci_one_batch_and_waves(built_lib, oracle_c):
    """Several automata (different kernel widths) in one call, and a workspace so small that
    the batch needs several waves."""
    from warpstr_b200.caller import CallerEngine
    eng = CallerEngine(workspace_bytes=6 << 20)
    sigs, aut, want = [], [], []
    for name in ('HD', 'CAN', 'AAAT'):
        locus, stas, ids, reads = _setup(eng, name, 5, seed=33)
        for r in reads:
            sigs.append(r.signal); aut.append(ids[int(r.reverse)])
            tb = co.tables_from(stas[int(r.reverse)])
            want.append(oracle_c.warp(r.signal, tb, np.zeros(len(r.signal), dtype=bool), 4, 110))
    traces = eng.warp_batch(sigs, aut)
    for t, w in zip(traces, want):
        assert np.array_equal(t, w)


def test_unreachable_end_and_tiny_reads(engine, oracle_c):
    """A read far too short to reach the end state: the reference's traceback stays in the
    end state all the way (every delta is inf/nan)."""
    locus, stas, ids, reads = _setup(engine, 'AAAT', 1, seed=4)
    sig = reads[0].signal[:300].copy()
    a = ids[int(reads[0].reverse)]
    tb = co.tables_from(stas[int(reads[0].reverse)])
    t = engine.warp_batch([sig, sig[:5], sig[:37]], [a, a, a])
    for s, got in zip([sig, sig[:5], sig[:37]], t):
        want = oracle_c.warp(s, tb, np.zeros(len(s), dtype=bool), 4, 110)
        assert np.array_equal(got, want)
    with pytest.raises(IndexError):
        engine.warp_batch([sig[:4]], [a])


def test_full_call_matches_oracle(engine, oracle_c):
    """Two passes + mid-stage: allele lengths bit-exact, costs to 1e-9 relative (the host
    mid-stage makes the same numpy/scipy calls, so they are in fact identical)."""
    for name in ('HD', 'FMR1'):
        locus, stas, ids, reads = _setup(engine, name, 5, seed=12)
        res = engine.call_batch([r.signal for r in reads], [ids[int(r.reverse)] for r in reads],
                                [r.reverse for r in reads])
        for r, got in zip(reads, res):
            tb = co.tables_from(stas[int(r.reverse)])
            want = co.run_read(r.signal, tb, 110, r.reverse, impl='c')
            assert got.seq == want.seq and got.resc_seq == want.resc_seq
            assert got.cost == pytest.approx(want.cost, rel=1e-9)
            assert got.resc_cost == pytest.approx(want.resc_cost, rel=1e-9)


@pytest.mark.parametrize('name', ['AAAT', 'HD', 'FMR1', 'DM2', 'CAN'])
def test_device_call_intermediates_match_oracle(engine, oracle_c, name):
    """wstr_call_batch end to end on the device: both traces identical, the rescaled signal
    bit-identical to scipy's splev of splrep(s=m), lengths / sequences exact, costs bit-equal
    (numpy pairwise-sum order is reproduced)."""
    locus, stas, ids, reads = _setup(engine, name, 6, seed=17, noise=0.2)
    sigs = [r.signal for r in reads]
    aut = [ids[int(r.reverse)] for r in reads]
    rev = [r.reverse for r in reads]
    packed = engine.upload(sigs, aut, rev)
    o = engine.call_packed(*packed, want_debug=True)
    off, lengths = packed[1], packed[2]
    assert not o['status'].cpu().numpy().any()
    t1, t2, resc = o['trace1'].cpu().numpy(), o['trace2'].cpu().numpy(), o['rescaled'].cpu().numpy()
    res = engine.results_from(o, sigs, aut, rev)
    for n, r in enumerate(reads):
        tb = co.tables_from(stas[int(r.reverse)])
        want = co.run_read(r.signal, tb, 110, r.reverse, impl='c')
        a, ln = int(off[n]), int(lengths[n])
        assert np.array_equal(t1[a:a + ln], want.trace1), (name, n)
        assert np.array_equal(resc[a:a + ln], want.rescaled), (name, n)
        assert np.array_equal(t2[a:a + ln], want.trace2), (name, n)
        assert res[n].seq == want.seq and res[n].resc_seq == want.resc_seq
        assert res[n].cost == want.cost and res[n].resc_cost == want.resc_cost


def test_untame_samples_take_the_guarded_divisions(engine, oracle_c):
    """The mid-stage divides by repeated divisors with a 5-instruction exact sequence when the
    operands are 'tame' (0 or within 2^+-120) and with the plain division otherwise.  Samples far
    outside that range (tiny ones inside the read, huge ones at its end, where they cannot upset
    the alignment) must give scipy's bits too: rescaled signal, both traces, costs."""
    locus, stas, ids, reads = _setup(engine, 'HD', 4, seed=31, noise=0.2)
    rng = np.random.default_rng(5)
    sigs = []
    for r in reads:
        x = r.signal.copy()
        idx = rng.integers(200, len(x) - 200, 12)
        x[idx[:6]] = 2.0 ** -130
        x[idx[6:9]] = -2.0 ** -140
        x[idx[9:]] = 0.0
        x[-1] = 2.0 ** 130
        x[-2] = -2.0 ** 125
        sigs.append(x)
    aut = [ids[int(r.reverse)] for r in reads]
    rev = [r.reverse for r in reads]
    packed = engine.upload(sigs, aut, rev)
    o = engine.call_packed(*packed, want_debug=True)
    off, lengths = packed[1], packed[2]
    assert not o['status'].cpu().numpy().any()
    t1, t2, resc = o['trace1'].cpu().numpy(), o['trace2'].cpu().numpy(), o['rescaled'].cpu().numpy()
    res = engine.results_from(o, sigs, aut, rev)
    for n, r in enumerate(reads):
        tb = co.tables_from(stas[int(r.reverse)])
        want = co.run_read(sigs[n], tb, 110, r.reverse, impl='c')
        a, ln = int(off[n]), int(lengths[n])
        assert np.array_equal(t1[a:a + ln], want.trace1), n
        assert np.array_equal(resc[a:a + ln], want.rescaled), n
        assert np.array_equal(t2[a:a + ln], want.trace2), n
        assert res[n].seq == want.seq and res[n].resc_seq == want.resc_seq
        assert res[n].cost == want.cost and res[n].resc_cost == want.resc_cost
        assert len(res[n].resc_seq) == r.truth_len


def test_device_and_host_engines_agree(engine):
    locus, stas, ids, reads = _setup(engine, 'HD', 24, seed=29, noise=0.3)
    args = ([r.signal for r in reads], [ids[int(r.reverse)] for r in reads], [r.reverse for r in reads])
    dev = engine.call_batch(*args, engine='gpu')
    host = engine.call_batch(*args, engine='host')
    for d, h in zip(dev, host):
        assert d.seq == h.seq and d.resc_seq == h.resc_seq
        assert d.cost == pytest.approx(h.cost, rel=1e-12) and d.resc_cost == pytest.approx(h.resc_cost, rel=1e-12)


def test_call_arrays_pipelined_chunks_equal_single_call(engine):
    """The end-to-end array API (chunked, copy/compute overlap) returns exactly what one
    device-resident call returns."""
    import torch
    locus, stas, ids, reads = _setup(engine, 'HD', 40, seed=61)
    sigs = [r.signal for r in reads]
    aut = np.array([ids[int(r.reverse)] for r in reads], dtype=np.int32)
    rev = np.array([r.reverse for r in reads], dtype=np.uint8)
    from warpstr_b200.caller import pack_signals
    host, off, lengths = pack_signals(sigs)
    one = engine.call_packed(host.cuda(), off, lengths, aut, rev)
    for chunk in (7, 40, 1000):
        got = engine.call_arrays(host, off, lengths, aut, rev, chunk_reads=chunk)
        for k in ('len1', 'len2', 'cost1', 'cost2', 'status'):
            assert np.array_equal(got[k], one[k].cpu().numpy()), (k, chunk)
        s1 = one['seq2'].cpu().numpy()
        for r in range(len(reads)):
            a, b = int(got['seq_off'][r]), int(one['seq_off'][r])
            n = int(got['len2'][r])
            assert np.array_equal(got['seq2'][a:a + n], s1[b:b + n])


@pytest.mark.parametrize('name,flank_arg', [('HD', 120), ('AAAT', 121), ('DM2', 123), ('CAN', 122)])
def test_open_end_band(engine, oracle_c, name, flank_arg):
    """A flank_length a little larger than the real right flank puts the end band's cut inside
    the repeat region (seq_idx counts regex positions, the loop body once), where back edges
    lead from kept states into skipped ones: the kernel must then keep the per-cell band test
    for every banded row (DevAutomaton.band_closed == 0)."""
    locus = synth.make_locus(name, seed=31, flank_length=110)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    ids = [engine.add_automaton(s, flank_arg) for s in stas]
    assert all(engine.automata[i].info()['band_closed'] == 0 for i in ids)
    assert engine.automata[engine.add_automaton(stas[0], 110)].info()['band_closed'] == 1
    reads = synth.make_reads(locus, 5, seed=131)
    sigs = [r.signal for r in reads]
    aut = [ids[int(r.reverse)] for r in reads]
    traces, costs = engine.warp_batch(sigs, aut, return_end_cost=True)
    rng = np.random.default_rng(8)
    masks = []
    for r in reads:
        m = np.zeros(len(r.signal), dtype=bool)
        m[rng.integers(0, len(m), 200)] = True
        m[len(m) - 700:len(m) - 500] = True          # a masked stretch across the band's first rows
        masks.append(m)
    traces_m = engine.warp_batch(sigs, aut, masks)
    for r, t, c, tm, m in zip(reads, traces, costs, traces_m, masks):
        tb = co.tables_from(stas[int(r.reverse)])
        m0 = np.zeros(len(r.signal), dtype=bool)
        assert np.array_equal(t, oracle_c.warp(r.signal, tb, m0, 4, flank_arg)), (name, r.name)
        assert c == oracle_c.fill(r.signal, tb, m0, 4, flank_arg)[-1, tb.endstate]
        assert np.array_equal(tm, oracle_c.warp(r.signal, tb, m, 4, flank_arg)), (name, r.name, 'masked')


def test_band_starts_on_any_row_phase(engine, oracle_c):
    """Reads whose lengths differ by one sample put the band's first row (and the last, partly
    filled direction word) on every phase of the 3-row cycle."""
    locus, stas, ids, reads = _setup(engine, 'HD', 2, seed=77)
    base = reads[0]
    sigs = [np.ascontiguousarray(base.signal[:len(base.signal) - k]) for k in range(7)]
    aut = [ids[int(base.reverse)]] * len(sigs)
    traces, costs = engine.warp_batch(sigs, aut, return_end_cost=True)
    tb = co.tables_from(stas[int(base.reverse)])
    for x, t, c in zip(sigs, traces, costs):
        m0 = np.zeros(len(x), dtype=bool)
        assert np.array_equal(t, oracle_c.warp(x, tb, m0, 4, 110)), len(x)
        assert c == oracle_c.fill(x, tb, m0, 4, 110)[-1, tb.endstate]
