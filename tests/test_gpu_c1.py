"""GPU: BASELINE config 1 -- the reference's bundled caller-only test (10 reads of the (AAAT)
locus Human_STR_1108232), raw int16 reads -> normalisation kernel -> two-pass caller.  The
fixture (scripts/make_c1_fixture.py) holds the raw reads, the pile-up-consensus flanks and
the oracle's answers; the reference's README.md:55 gives the genotype (44, 40)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _fixture():
    z = np.load(os.path.join(GOLD, 'c1_bundled.npz'))
    raws = [np.cumsum(z[f'raw_delta{i}'].astype(np.int64)).astype(np.int16) for i in range(len(z['names']))]
    wins = list(zip(z['l_start_raw'].tolist(), z['r_end_raw'].tolist()))
    return z, raws, wins


def test_bundled_reads_normalise_like_numpy(built_lib):
    from oracle import normalize_oracle as no
    from warpstr_b200.normalize import normalize_windows
    z, raws, wins = _fixture()
    got, ss = normalize_windows(raws, wins, 'Brute', return_shift_scale=True)
    for i, (raw, win) in enumerate(zip(raws, wins)):
        want = no.get_data_processed(raw, win)
        assert np.array_equal(got[i], want), i            # bit-exact float64
        assert np.array_equal(ss[i], z['shift_scale'][i]), i


def test_bundled_caller_only_genotype(built_lib):
    from warpstr_b200.normalize import normalize_windows
    from warpstr_b200.wrapper import CallerWrapper, Locus, ReadSignal, flanks_from_template
    z, raws, wins = _fixture()
    sigs = normalize_windows(raws, wins, 'Brute')
    workload = [ReadSignal(str(n), bool(r), s) for n, r, s in zip(z['names'], z['reverse'], sigs)]
    locus = Locus(name='Human_STR_1108232', sequence=str(z['sequence']), flank_length=int(z['flank_length']))
    cw = CallerWrapper(locus, threads=2, flanks=flanks_from_template(str(z['left']), str(z['right'])))
    res = cw.run(workload)
    assert [len(r.seq) for r in res] == z['len1'].tolist()
    assert [len(r.resc_seq) for r in res] == z['len2'].tolist()
    assert [r.resc_seq for r in res] == [str(s) for s in z['resc_seq']]
    assert [r.seq for r in res] == [str(s) for s in z['seq']]
    np.testing.assert_allclose([r.cost for r in res], z['cost1'], rtol=1e-9)
    np.testing.assert_allclose([r.resc_cost for r in res], z['cost2'], rtol=1e-9)
    # the two alleles the reference's README reports for this test
    assert sorted(set(len(r.resc_seq) for r in res)) == [40, 44]


def test_bundled_reads_device_resident_chain(built_lib):
    """CallerWrapper.run_raw: the same ten reads as DAC counts in, calls out -- normalisation kernel and
    caller chained on the device (wrapper.get_workload + run, wrapper.py:44-54,104-120), the float64
    windows never on the host.  Same sequences, lengths and costs as the fixture."""
    from warpstr_b200.wrapper import CallerWrapper, Locus, flanks_from_template
    z, raws, wins = _fixture()
    locus = Locus(name='Human_STR_1108232', sequence=str(z['sequence']), flank_length=int(z['flank_length']))
    cw = CallerWrapper(locus, threads=2, flanks=flanks_from_template(str(z['left']), str(z['right'])))
    res = cw.run_raw(raws, wins, [bool(r) for r in z['reverse']])
    assert [r.resc_seq for r in res] == [str(s) for s in z['resc_seq']]
    assert [r.seq for r in res] == [str(s) for s in z['seq']]
    assert [r.cost for r in res] == z['cost1'].tolist()
    assert [r.resc_cost for r in res] == z['cost2'].tolist()
    assert sorted(set(len(r.resc_seq) for r in res)) == [40, 44]
