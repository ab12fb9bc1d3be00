"""GPU: the three forms a batch can be handed to the engine in -- float64 windows (the reference's
ReadSignal.signal), int16 window samples + {shift, scale} (wstr_dequantize_batch), raw int16 reads +
windows (wstr_normalize_batch) -- give the same results bit for bit, and the oracle's."""
import numpy as np
import pytest

from oracle import caller_oracle as co
from oracle import normalize_oracle as no
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata

pytestmark = pytest.mark.gpu


def _batch(n, seed, name='HD'):
    locus = synth.make_locus(name, seed=seed)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    reads = synth.make_reads(locus, n, seed=seed + 1, noise=0.2)
    return locus, stas, reads


def test_dequantize_is_the_reference_expression(built_lib):
    """(raw - shift) / scale on the device == numpy on the host, every bit, aligned or not."""
    import torch
    from warpstr_b200 import _lib
    rng = np.random.default_rng(2)
    lens = [1, 7, 8, 9, 1000, 3333, 16]
    raw_off = np.array([0, 8, 16, 24, 40, 1048, 4392], dtype=np.int64)          # last one deliberately unaligned below
    raw_off[-1] += 3
    total = int(raw_off[-1] + lens[-1])
    raw = rng.integers(-300, 1500, size=total).astype(np.int16)
    ss = np.stack((rng.normal(530, 30, len(lens)), rng.normal(57, 5, len(lens))), axis=1)
    out_off = np.array([0, 2, 10, 18, 28, 1028, 4363], dtype=np.int64)           # last start odd: scalar path
    d_out = torch.zeros(int(out_off[-1] + lens[-1]) + 2, dtype=torch.float64, device='cuda')
    ws = torch.empty(24 * len(lens) + 256, dtype=torch.uint8, device='cuda')
    _lib.dequantize_batch(torch.from_numpy(raw).cuda(), raw_off, np.array(lens, dtype=np.int32),
                          torch.from_numpy(ss.reshape(-1)).cuda(), d_out, out_off, ws)
    got = d_out.cpu().numpy()
    for r, n in enumerate(lens):
        want = (raw[raw_off[r]:raw_off[r] + n] - ss[r, 0]) / ss[r, 1]             # schemas/fast5.py:113
        assert np.array_equal(got[out_off[r]:out_off[r] + n], want), r


def test_three_ingestion_forms_agree(built_lib, oracle_c):
    import torch
    from warpstr_b200.caller import CallerEngine, pack_signals
    locus, stas, reads = _batch(40, seed=11)
    eng = CallerEngine()
    ids = [eng.add_automaton(s, 110) for s in stas]
    aut = np.array([ids[int(r.reverse)] for r in reads], dtype=np.int32)
    rev = np.array([r.reverse for r in reads], dtype=np.uint8)
    rng = np.random.default_rng(5)

    # raw reads (window embedded in a longer read, spikes injected) -> the reference's normalisation on the host
    raws, wins = [], []
    for r in reads:
        raw, lo, hi = synth.to_raw_int16(rng, r.signal, pad=4096, spike_rate=5e-4)
        raws.append(raw)
        wins.append((lo, hi))
    patched = [no.remove_spikes(raw, 'Brute') for raw in raws]
    shift_scale = []
    for p in patched:
        shift = np.mean(np.percentile(p, (46.5, 53.5)))
        shift_scale.append((shift, np.median(np.abs(p - shift))))
    windows64 = [no.get_data_processed(raw, w, 'Brute') for raw, w in zip(raws, wins)]

    # (a) float64 windows
    host, off, lengths = pack_signals(windows64)
    a = eng.call_arrays(host, off, lengths, aut, rev, chunk_reads=16)
    a = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    # (b) int16 window samples + {shift, scale}
    lens16 = np.array([w[1] - w[0] + 1 for w in wins], dtype=np.int32)
    raw_off = np.zeros(len(reads), dtype=np.int64)
    raw_off[1:] = np.cumsum((lens16[:-1].astype(np.int64) + 7) & ~7)
    h16 = torch.zeros(int(raw_off[-1] + lens16[-1] + 8), dtype=torch.int16).pin_memory()
    for p, w, o in zip(patched, wins, raw_off):
        h16.numpy()[o:o + w[1] - w[0] + 1] = p[w[0]:w[1] + 1]
    hss = torch.from_numpy(np.array(shift_scale, dtype=np.float64)).pin_memory()
    b = eng.call_arrays_quantized(h16, raw_off, lens16, hss, aut, rev, chunk_reads=16)
    b = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    # (c) raw reads + windows
    roff = np.zeros(len(reads) + 1, dtype=np.int64)
    roff[1:] = np.cumsum([len(x) for x in raws])
    hraw = torch.from_numpy(np.concatenate(raws)).pin_memory()
    c = eng.call_arrays_raw(hraw, roff, [w[0] for w in wins], [w[1] for w in wins], aut, rev, 'Brute', chunk_reads=16)

    assert not a['status'].any()
    for other in (b, c):
        for k in ('len1', 'len2', 'cost1', 'cost2', 'status', 'ttest_ties'):
            assert np.array_equal(a[k], other[k]), k
        for r in range(len(reads)):
            s0 = int(a['seq_off'][r])
            assert np.array_equal(a['seq2'][s0:s0 + a['len2'][r]], other['seq2'][s0:s0 + a['len2'][r]])
    assert np.array_equal(c['shift_scale'], np.array(shift_scale))
    # and the oracle on the float64 windows
    for r in (0, 7, 23, 39):
        want = co.run_read(windows64[r], co.tables_from(stas[int(rev[r])]), 110, bool(rev[r]), impl='c')
        s0 = int(a['seq_off'][r])
        assert a['seq2'][s0:s0 + a['len2'][r]].tobytes().decode() == want.resc_seq
        assert a['cost2'][r] == want.resc_cost and a['cost1'][r] == want.cost


def test_edge_inputs_of_the_array_entry_points(built_lib):
    """A window that runs past the end of its read (numpy clips the slice), an empty window, a window shorter
    than the dwell, empty batches."""
    import torch
    from warpstr_b200.caller import CallerEngine
    locus, stas, reads = _batch(4, seed=21)
    eng = CallerEngine()
    ids = [eng.add_automaton(s, 110) for s in stas]
    rng = np.random.default_rng(0)
    raws, wins = [], []
    for r in reads:
        raw, lo, hi = synth.to_raw_int16(rng, r.signal, pad=512)
        raws.append(raw)
        wins.append((lo, hi))
    wins[1] = (wins[1][0], len(raws[1]) + 1000)
    wins[2] = (100, 50)
    wins[3] = (10, 12)
    roff = np.zeros(5, dtype=np.int64)
    roff[1:] = np.cumsum([len(x) for x in raws])
    aut = np.array([ids[int(r.reverse)] for r in reads], dtype=np.int32)
    rev = np.array([r.reverse for r in reads], dtype=np.uint8)
    res = eng.call_arrays_raw(torch.from_numpy(np.concatenate(raws)).pin_memory(), roff, [w[0] for w in wins],
                              [w[1] for w in wins], aut, rev)
    assert res['status'].tolist() == [0, 0, 1, 1]
    assert res['lengths'].tolist() == [len(reads[0].signal), len(raws[1]) - wins[1][0], 0, 3]
    assert res['len2'][2] == -1 and np.isnan(res['cost2'][3]) and res['len2'][0] > 0
    want = no.get_data_processed(raws[1], wins[1], 'Brute')
    x1 = co.run_read(want, co.tables_from(stas[int(rev[1])]), 110, bool(rev[1]), impl='c')
    assert res['len2'][1] == len(x1.resc_seq) and res['cost2'][1] == x1.resc_cost
    with pytest.raises(IndexError):                     # the reference: signal[0] / D[0, i] on an empty window
        eng.call_raw_batch(raws, wins, [int(a) for a in aut], [bool(x) for x in rev])
    e = np.zeros(0)
    pin = lambda t: t.pin_memory()
    assert eng.call_arrays(pin(torch.zeros(2, dtype=torch.float64)), e.astype(np.int64), e.astype(np.int32),
                           e.astype(np.int32), e.astype(np.uint8))['len1'].shape == (0,)
    assert eng.call_arrays_quantized(pin(torch.zeros(8, dtype=torch.int16)), e.astype(np.int64), e.astype(np.int32),
                                     pin(torch.zeros((0, 2), dtype=torch.float64)), e.astype(np.int32),
                                     e.astype(np.uint8))['len1'].shape == (0,)
    assert eng.call_batch([], [], []) == [] and eng.call_raw_batch([], [], [], []) == []
