"""GPU: the CUDA path against the committed reference goldens (tests/golden/, made from the
unmodified reference by oracle/make_golden.py) -- the whole call, the normalisation kernel
and the pore-model lookup, all through the C ABI."""
import json
import os

import numpy as np
import pytest

from warpstr_b200.automata import StateAutomata

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_whole_call_matches_reference_goldens(built_lib):
    from warpstr_b200.caller import CallerEngine
    z = np.load(os.path.join(GOLD, 'caller.npz'))
    cases = json.loads(str(z['cases']))
    by_flank = {}
    for c in cases:
        by_flank.setdefault(c['flank'], []).append(c)
    for flank, group in by_flank.items():
        eng = CallerEngine()
        ids = {}
        sigs, aut, rev = [], [], []
        for c in group:
            rx = c['reverse_regex'] if c['reverse'] else c['template_regex']
            if rx not in ids:
                ids[rx] = eng.add_automaton(StateAutomata(rx), flank)
            sigs.append(z[f"{c['key']}_signal"]); aut.append(ids[rx]); rev.append(c['reverse'])
        packed = eng.upload(sigs, aut, rev)
        o = eng.call_packed(*packed, want_debug=True)
        assert not o['status'].cpu().numpy().any()
        t1, t2, resc = (o[k].cpu().numpy() for k in ('trace1', 'trace2', 'rescaled'))
        res = eng.results_from(o, sigs, aut, rev)
        for n, c in enumerate(group):
            k = c['key']
            a, ln = int(packed[1][n]), int(packed[2][n])
            assert np.array_equal(t1[a:a + ln], z[f'{k}_ref_trace1']), k
            assert np.array_equal(resc[a:a + ln], z[f'{k}_ref_rescaled']), k
            assert np.array_equal(t2[a:a + ln], z[f'{k}_ref_trace2']), k
            assert res[n].seq == c['seq'] and res[n].resc_seq == c['resc_seq'], k
            assert len(res[n].resc_seq) == c['truth_len'], k
            # float tolerance of the north star: 1e-6 relative; in fact bit-equal
            assert res[n].cost == pytest.approx(c['cost'], rel=1e-6) and res[n].cost == c['cost'], k
            assert res[n].resc_cost == c['resc_cost'], k


@pytest.mark.parametrize('mode', ['Brute', 'None', 'median3', 'median5'])
def test_normalisation_kernel_matches_reference_goldens(built_lib, mode):
    from warpstr_b200.normalize import normalize_windows
    z = np.load(os.path.join(GOLD, 'normalize.npz'))
    cases = json.loads(str(z['cases']))
    key = {'Brute': 'brute', 'None': 'none', 'median3': 'median3', 'median5': 'median5'}[mode]
    use = [c for c in cases if f"{c['key']}_ref_norm_{key}" in z]
    raws = [z[f"{c['key']}_raw"] for c in use]
    wins = [(c['lo'], c['hi']) for c in use]
    got = normalize_windows(raws, wins, mode)
    for c, g in zip(use, got):
        want = z[f"{c['key']}_ref_norm_{key}"]
        assert g.shape == want.shape, c['key']
        assert np.array_equal(g, want, equal_nan=True), (c['key'], mode)


def test_normalisation_window_clipping_and_many_reads(built_lib):
    from oracle import normalize_oracle as no
    from warpstr_b200.normalize import get_data_processed, normalize_windows
    rng = np.random.default_rng(8)
    raws = [rng.integers(250, 1000, size=int(n)).astype(np.int16) for n in rng.integers(3000, 9000, size=400)]
    for r in raws[::7]:
        r[rng.integers(0, len(r), 5)] = 1700
    wins = [(int(rng.integers(0, len(r) - 10)), int(rng.integers(0, 2 * len(r)))) for r in raws]
    got = normalize_windows(raws, wins, 'Brute')
    for r, w, g in zip(raws, wins, got):
        assert np.array_equal(g, no.get_data_processed(r, w, 'Brute'))
    assert np.array_equal(get_data_processed(raws[0]), no.get_data_processed(raws[0], (0, len(raws[0]) - 1)))


@pytest.mark.parametrize('mode', ['None', 'Brute', 'median3', 'median5'])
def test_normalisation_edges(built_lib, mode):
    """Reads of every small length (all eight alignments of the first sample inside a 16-byte
    vector, reads shorter than a tile, tile-boundary lengths), spikes on the first / last samples
    and across tile boundaries, out-of-histogram values, clipped and empty windows."""
    from oracle import normalize_oracle as no
    from warpstr_b200.normalize import normalize_windows
    rng = np.random.default_rng(int.from_bytes(mode.encode(), 'little') % 1000)
    lens = list(range(1, 40)) + [2047, 2048, 2049, 4095, 4096, 4097, 6143, 6150] + \
        [int(v) for v in rng.integers(40, 7000, size=60)]
    raws, wins = [], []
    for n in lens:
        r = rng.integers(300, 900, size=n).astype(np.int16)
        for pos in (0, 1, 2, 3, n - 3, n - 2, n - 1, 2046, 2047, 2048, 2049, 4095, 4096):
            if 0 <= pos < n and rng.random() < 0.6:
                r[pos] = rng.choice([1500, 100, 1001, 249, 20000, -700])
        if n > 50 and rng.random() < 0.5:                      # a run of adjacent spikes
            a = int(rng.integers(3, n - 10))
            r[a:a + 4] = 1800
        raws.append(r)
        lo = int(rng.integers(0, n))
        wins.append((lo, int(rng.integers(lo - 1 if lo else 0, 2 * n))))
    got = normalize_windows(raws, wins, mode)
    for r, w, g in zip(raws, wins, got):
        want = no.get_data_processed(r, w, mode)
        assert g.shape == want.shape, (len(r), w)
        assert np.array_equal(g, want, equal_nan=True), (len(r), w, mode)


@pytest.mark.parametrize('mode', ['None', 'Brute'])
def test_normalisation_window_histogram_and_its_fallback(built_lib, mode):
    """Brute / None count the samples in a 512-value window around an estimate of the read's median
    (one counter column per lane); a read whose order statistics fall outside that window is redone on
    the value-indexed histogram.  Narrow, wide, bimodal, constant, drifting and off-scale reads, with
    and without spikes: all bit-equal to numpy."""
    from oracle import normalize_oracle as no
    from warpstr_b200.normalize import normalize_windows
    rng = np.random.default_rng(31 + len(mode))
    raws = []
    for n in (37, 2048, 5000, 40000, 150001):
        t = np.arange(n)
        raws += [
            (470 + 28 * rng.standard_normal(n)).astype(np.int16),                       # a real read's spread
            (500 + 110 * rng.standard_normal(n)).astype(np.int16),                      # tails outside the window
            np.where(rng.random(n) < 0.5, 260, 940).astype(np.int16),                   # bimodal: MAD outside it
            np.where(rng.random(n) < 0.48, 300 + rng.integers(0, 5, n), 800).astype(np.int16),
            np.full(n, 512, dtype=np.int16),                                            # scale 0 -> inf / nan
            (300 + 500 * t / n + 10 * rng.standard_normal(n)).astype(np.int16),         # drift across the read
            (9000 + 30 * rng.standard_normal(n)).astype(np.int16),                      # beyond the 8192-bin histogram
            (-200 + 30 * rng.standard_normal(n)).astype(np.int16),                      # negative
            np.concatenate((np.full(n // 2, 300), np.full(n - n // 2, 811))).astype(np.int16),   # estimate far off
            rng.integers(-32768, 32767, size=n).astype(np.int16),                       # the whole int16 range
        ]
    for i, r in enumerate(raws):
        if i % 2 == 0 and len(r) > 100:
            r[rng.integers(3, len(r), max(len(r) // 3000, 3))] = rng.choice([1500, 90, 30000])
    raws.append(np.full(2_200_000, 480, dtype=np.int16))                                 # too long for the 16-bit counters
    raws.append(np.full(2_090_000, 481, dtype=np.int16))                                 # ... and nearly: one bin takes it all
    r = (470 + 28 * rng.standard_normal(30000)).astype(np.int16)                        # more spikes than the list holds
    r[rng.integers(3, len(r), 600)] = 1400
    raws.append(r)
    wins = [(int(rng.integers(0, max(len(r) - 3000, 1))), 0) for r in raws]
    wins = [(lo, lo + 2999) for lo, _ in wins]
    with np.errstate(all='ignore'):
        got = normalize_windows(raws, wins, mode, return_shift_scale=True)
        for i, (r, w) in enumerate(zip(raws, wins)):
            want = no.get_data_processed(r, w, mode)
            assert np.array_equal(got[0][i], want, equal_nan=True), (i, len(r), mode)


def test_pore_lookup_edges(built_lib):
    """Every length around the 4-outputs-per-thread grouping, invalid characters at every offset,
    an output buffer that is not 16-byte aligned."""
    import torch
    from warpstr_b200 import _lib
    from warpstr_b200.pore_model import get_pore_model
    pm = get_pore_model()
    table = np.ascontiguousarray(pm.table)
    d_tab = torch.from_numpy(table).cuda()
    rng = np.random.default_rng(12)
    code = {c: i for i, c in enumerate('ACGT')}
    for n in list(range(6, 40)) + [1023, 1024, 1029, 5000]:
        seq = ''.join(rng.choice(list('ACGT'), size=n))
        for bad_at in (None, 0, 3, n - 1, n // 2):
            s = seq if bad_at is None else seq[:bad_at] + 'N' + seq[bad_at + 1:]
            want = np.full(n - 5, np.nan)
            for i in range(n - 5):
                kmer = s[i:i + 6]
                if 'N' not in kmer:
                    idx = 0
                    for ch in kmer:
                        idx = idx * 4 + code[ch]
                    want[i] = table[idx]
            d_seq = torch.from_numpy(np.frombuffer(s.encode(), dtype=np.uint8).copy()).cuda()
            for shift in (0, 1):                               # shift 1: output not 16-byte aligned
                buf = torch.full((n - 5 + 2,), -1.0, dtype=torch.float64, device='cuda')
                d_bad = torch.zeros(1, dtype=torch.int32, device='cuda')
                _lib.pore_lookup(d_seq, d_tab, 6, buf[shift:shift + n - 5], d_bad)
                got = buf.cpu().numpy()
                assert np.array_equal(got[shift:shift + n - 5], want, equal_nan=True), (n, bad_at, shift)
                assert got[shift + n - 5] == -1.0 and (shift == 0 or got[0] == -1.0)
                assert int(d_bad.item()) == int(np.isnan(want).sum())


def test_pore_lookup_long_sequences(built_lib):
    """Sequences of 2^18 levels and more take the shared-memory-table kernel (8 levels per thread, 8-byte
    base loads); same contract: invalid characters anywhere give NaN for every k-mer that holds them and
    are counted, ragged ends, and unaligned inputs fall back to the plain kernel with the same result."""
    import torch
    from warpstr_b200 import _lib
    from warpstr_b200.pore_model import get_pore_model
    table = np.ascontiguousarray(get_pore_model().table)
    d_tab = torch.from_numpy(table).cuda()
    rng = np.random.default_rng(21)
    for n in (262149, 300003, 524301):
        codes = rng.integers(0, 4, n)
        seq = np.frombuffer(b'ACGT', dtype=np.uint8)[codes].copy()
        bad_at = np.array([0, 5, 6, 1000, 1001, 4095, 4096, n // 2, n - 12, n - 6, n - 1])
        seq[bad_at] = np.frombuffer(b'NaNcgtN-NxN', dtype=np.uint8)
        valid = np.ones(n, dtype=bool)
        valid[bad_at] = False
        idx = np.zeros(n - 5, dtype=np.int64)
        ok = np.ones(n - 5, dtype=bool)
        for p in range(6):
            idx = idx * 4 + codes[p:p + n - 5]
            ok &= valid[p:p + n - 5]
        want = np.where(ok, table[idx], np.nan)
        buf_seq = torch.zeros(n + 8, dtype=torch.uint8, device='cuda')
        for s_shift, o_shift in ((0, 0), (1, 0), (0, 1)):
            buf_seq[s_shift:s_shift + n] = torch.from_numpy(seq).cuda()
            buf = torch.full((n - 5 + 2,), -1.0, dtype=torch.float64, device='cuda')
            d_bad = torch.zeros(1, dtype=torch.int32, device='cuda')
            _lib.pore_lookup(buf_seq[s_shift:s_shift + n], d_tab, 6, buf[o_shift:o_shift + n - 5], d_bad)
            got = buf.cpu().numpy()
            assert np.array_equal(got[o_shift:o_shift + n - 5], want, equal_nan=True), (n, s_shift, o_shift)
            assert got[o_shift + n - 5] == -1.0 and (o_shift == 0 or got[0] == -1.0)
            assert int(d_bad.item()) == int((~ok).sum())


def test_pore_lookup_matches_reference_golden(built_lib):
    from warpstr_b200.pore_model import get_pore_model
    sq = np.load(os.path.join(GOLD, 'squiggle.npz'))
    pm = get_pore_model()
    assert np.array_equal(pm.generate_signal(str(sq['seq'])), sq['ref_signal'])
    assert pm.generate_signal('ACG').shape == (0,)
    with pytest.raises(IndexError):
        pm.generate_signal('ACGTNACGTAC')


def test_engine_raises_reference_exception_types(built_lib):
    """A read the reference would abort on is reported with the reference's exception type."""
    from warpstr_b200 import synth
    from warpstr_b200.caller import CallerEngine
    locus = synth.make_locus('AAAT', seed=2)
    eng = CallerEngine()
    a = eng.add_automaton(StateAutomata(locus.template_regex), 110)
    rd = [r for r in synth.make_reads(locus, 4, seed=3) if not r.reverse][0]
    with pytest.raises(TypeError):       # never reaches the repeat: splrep has < 4 points
        eng.call_batch([rd.signal[:300]], [a], [False])
    with pytest.raises(IndexError):      # T <= min_values_per_state
        eng.call_batch([rd.signal[:3]], [a], [False])
    ok = eng.call_batch([rd.signal], [a], [False])
    assert len(ok[0].resc_seq) == rd.truth_len


def test_main_wrapper_files(built_lib, tmp_path):
    """overview.csv in -> calls on the GPU -> overview.csv / FASTA / complex-unit CSV out, through
    the reference-shaped driver (wrapper.py:17-41), for a complex (multi-unit) locus."""
    import pandas as pd
    from oracle import caller_oracle as co
    from warpstr_b200 import synth
    from warpstr_b200.wrapper import Locus, ReadSignal, flanks_from_template, main_wrapper
    sl = synth.make_locus('DM2', seed=8)
    reads = synth.make_reads(sl, 6, seed=9)
    d = tmp_path / 'DM2'
    (d / 'expected_signals').mkdir(parents=True)
    (d / 'summaries').mkdir()
    fl = flanks_from_template(sl.left, sl.right)
    with open(d / 'expected_signals' / 'sequences.csv', 'w') as fh:
        fh.write('type,sequence\n')
        fh.write(f'left_flank_template,{fl.template.left}\nright_flank_template,{fl.template.right}\n')
        fh.write(f'left_flank_reverse,{fl.reverse.left}\nright_flank_reverse,{fl.reverse.right}\n')
    pd.DataFrame({'read_name': [r.name for r in reads] + ['skipped'], 'run_id': ['x'] * 7,
                  'reverse': [r.reverse for r in reads] + [False], 'saved': [1] * 6 + [0],
                  'l_start_raw': [0] * 7, 'r_end_raw': [1] * 7}).to_csv(d / 'overview.csv', index=False)
    locus = Locus('DM2', sl.sequence, 110, str(d))
    df, dfc = main_wrapper(locus, 4, workload=[ReadSignal(r.name, r.reverse, r.signal) for r in reads])
    assert list(df['results'][:6]) == [r.truth_len for r in reads] and df['results'].iloc[6] == -1
    assert dfc is not None and 'main_CAGG' in dfc.columns and len(dfc) == 6
    assert os.path.exists(d / 'predictions' / 'sequences' / 'all.fasta')
    assert os.path.exists(d / 'summaries' / 'state_similarity.csv')
    back = pd.read_csv(d / 'overview.csv')
    assert {'results', 'orig', 'dtw_cost1', 'dtw_cost2'} <= set(back.columns)


def test_expected_signal_step_files(built_lib, tmp_path):
    """sequences.csv + expected-signal text files (Squiggler.process_locus), read back by load_flanks."""
    from warpstr_b200.locus import write_expected_signals
    from warpstr_b200.pore_model import get_pore_model
    from warpstr_b200.wrapper import load_flanks
    rng = np.random.default_rng(3)
    chrom = ''.join(rng.choice(list('ACGT'), 400)) + 'CAG' * 20 + ''.join(rng.choice(list('ACGT'), 400))
    fa = tmp_path / 'g.fa'
    fa.write_text('>chr9\n' + '\n'.join(chrom[i:i + 50] for i in range(0, len(chrom), 50)) + '\n')
    seqs = write_expected_signals(str(tmp_path / 'L'), 'chr9:401-460', str(fa), 110)
    assert seqs['temp_ref_pattern'] == 'CAG' * 20 and seqs['left_flank_template'] == chrom[290:400]
    fl = load_flanks(str(tmp_path / 'L'))
    assert fl.template.right == chrom[460:570] and fl.reverse.left == seqs['left_flank_reverse']
    lv = np.loadtxt(tmp_path / 'L' / 'expected_signals' / 'left_flank_template.txt')
    pm = get_pore_model()
    want = pm.get_values([seqs['left_flank_template'][i:i + 6] for i in range(105)])
    assert lv.shape == (105,) and np.allclose(lv, want, atol=1e-6)
