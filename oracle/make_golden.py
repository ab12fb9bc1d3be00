"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Every array stored under a 'ref_' key was produced by the reference's own functions
(imported through oracle/refshim.py): StateAutomata, WarpSTR._calc_dtw_astates,
WarpSTR._backtracking, WarpSTR.run, rescale_signal, mask_bad_repeats,
normalize_signal_mad, Fast5.brute_remove, pore_model.get_value.  Inputs are the seeded
synthetic reads of warpstr_b200.synth.  The GPU box has no reference tree; the tests there
(and the CPU tests) compare against these files.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from warpstr_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

CALLER_CASES = [  # (name, recipe, flank, n_reads, noise)
    ('AAAT', 'AAAT', 110, 2, 0.15),
    ('HD', 'HD', 110, 2, 0.15),
    ('FMR1', 'FMR1', 110, 2, 0.2),
    ('FMR1_MGG', 'FMR1_MGG', 110, 1, 0.15),
    ('DM2', 'DM2', 110, 2, 0.15),
    ('CAN', 'CAN', 110, 1, 0.15),
    ('AAAT_F40', 'AAAT', 40, 2, 0.15),
]

AUTOMATA_PATTERNS = ['(AAAT)', '(AGC)AACAGCCGCCAC(CGC)', '((CGG){AGG})', '(MGG)', '(GGGGCC)',
                     '((CAGG){CAGM})(CAGA)(CA)', '(CAN)', '(AARRG)', '(A{C}G)', '(AC{GT}(TA))', '{ACG}(TTC)',
                     '(N)', '(RY)', 'AC{G}{T}CA(GA)', '((A)(C))', '(AT)N(GC)', '(A{CN}T)', '(A)']


def automaton_record(sta):
    st = sta.states
    ptr = np.zeros(len(st) + 1, dtype=np.int32)
    idx = []
    for i, s in enumerate(st):
        idx.extend(p.idx for p in s.incoming)
        ptr[i + 1] = len(idx)
    return dict(kmers=np.array([s.kmer for s in st]), values=np.array([s.value for s in st], dtype=np.float64),
                seq_idx=np.array([s.seq_idx for s in st], dtype=np.int32), in_ptr=ptr,
                in_idx=np.array(idx, dtype=np.int32), mask=np.array(sta.mask, dtype=bool),
                endstate=np.int32(sta.endstate), repstart=np.int32(sta.repstart), repend=np.int32(sta.repend))


def main():
    ref = refshim.load()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # ---- pore model -------------------------------------------------------------------------------
    tbl = ref.pore_model.table
    np.savez_compressed(os.path.join(OUT, 'pore_model.npz'),
                        ref_level_norm=tbl['level_norm'].values.astype(np.float64),
                        kmers=np.array(list(tbl['kmer'].values)))

    # ---- automata ---------------------------------------------------------------------------------
    auto = {}
    meta = []
    for n, pat in enumerate(AUTOMATA_PATTERNS):
        for F in (110, 12):
            lf, rf = synth.random_flank(rng, F), synth.random_flank(rng, F)
            seq = lf + pat + rf
            rec = automaton_record(ref.StateAutomata(seq))
            key = f'a{len(meta)}'
            meta.append(dict(key=key, pattern=pat, sequence=seq))
            for k, v in rec.items():
                auto[f'{key}_ref_{k}'] = v
    np.savez_compressed(os.path.join(OUT, 'automata.npz'), meta=json.dumps(meta), **auto)

    # ---- caller -----------------------------------------------------------------------------------
    out = {}
    cases = []
    for name, recipe, F, n_reads, noise in CALLER_CASES:
        locus = synth.make_locus(name, seed=41, flank_length=F, recipe=recipe)
        reads = synth.make_reads(locus, n_reads, seed=43, noise=noise)
        autos = {False: ref.StateAutomata(locus.template_regex), True: ref.StateAutomata(locus.reverse_regex)}
        for i, rd in enumerate(reads):
            sta = autos[rd.reverse]
            w = ref.WarpSTR(F, sta.states, sta.endstate, sta.mask, None, rd.reverse, rd.name)
            m0 = np.full(len(rd.signal), False)
            D = w._calc_dtw_astates(rd.signal, sta.states, m0)
            t1 = w._backtracking(D, sta.states, rd.signal, m0)
            wr = ref.caller.WarpResult(t1)
            al = wr.create_alignment(sta.states, rd.signal)
            resc = ref.caller.rescale_signal(rd.signal, al)
            start, end, bad = ref.caller.mask_bad_repeats(rd.signal, sta.mask, t1, wr.state_transitions)
            bad = np.asarray(bad, dtype=bool)
            D2 = w._calc_dtw_astates(resc, sta.states, bad)
            t2 = w._backtracking(D2, sta.states, resc, bad)
            res = w.run(rd.signal)
            key = f'{name}_{i}'
            cases.append(dict(key=key, locus=name, recipe=recipe, flank=F, reverse=bool(rd.reverse),
                              template_regex=locus.template_regex, reverse_regex=locus.reverse_regex,
                              seq=res.seq, resc_seq=res.resc_seq, cost=float(res.cost),
                              resc_cost=float(res.resc_cost), start=int(start), end=int(end),
                              truth_len=int(rd.truth_len)))
            out[f'{key}_signal'] = rd.signal
            out[f'{key}_ref_trace1'] = t1.astype(np.int16)
            out[f'{key}_ref_trace2'] = t2.astype(np.int16)
            out[f'{key}_ref_rescaled'] = np.asarray(resc, dtype=np.float64)
            out[f'{key}_ref_badmask'] = np.packbits(bad)
            out[f'{key}_ref_D1_last'] = D[-1].copy()
            out[f'{key}_ref_D2_last'] = D2[-1].copy()
            # a sparse sample of the first matrix: every 97th row
            out[f'{key}_ref_D1_rows'] = D[::97].copy()
            print(key, len(rd.signal), len(res.resc_seq), rd.truth_len)
    np.savez_compressed(os.path.join(OUT, 'caller.npz'), cases=json.dumps(cases), **out)

    # ---- normalisation ----------------------------------------------------------------------------
    norm = {}
    ncases = []
    locus = synth.make_locus('HD', seed=5)
    for i, rd in enumerate(synth.make_reads(locus, 4, seed=6)):
        raw, lo, hi = synth.to_raw_int16(rng, rd.signal, pad=6000 + 1000 * i, spike_rate=[1e-4, 2e-3, 0.0, 5e-2][i])
        if i == 1:
            raw[:3] = [1400, 90, 1200]          # spikes at i <= 2 are left alone (fast5.py:98-99)
            raw[-2:] = [1500, 60]               # shortened median windows at the array end
            raw[500:506] = [1500, 1500, 90, 1500, 90, 90]   # chained spikes: later medians see earlier fixes
        if i == 3:
            raw[1000:1010] = [-3000, 30000, -32768, 32767, 9000, 8192, -1, 0, 8191, 20000]
        ncases.append(dict(key=f'n{i}', lo=int(lo), hi=int(hi)))
        norm[f'n{i}_raw'] = raw
        fixed = ref.Fast5.brute_remove(raw)
        norm[f'n{i}_ref_brute'] = fixed
        norm[f'n{i}_ref_norm_brute'] = ref.normalize_signal_mad(fixed)[lo:hi + 1]
        norm[f'n{i}_ref_norm_none'] = ref.normalize_signal_mad(raw)[lo:hi + 1]
        from scipy.signal import medfilt
        norm[f'n{i}_ref_norm_median3'] = ref.normalize_signal_mad(medfilt(raw, 3))[lo:hi + 1]
        norm[f'n{i}_ref_norm_median5'] = ref.normalize_signal_mad(medfilt(raw, 5))[lo:hi + 1]
    # even / odd and tiny lengths
    for j, n in enumerate((1, 2, 5, 6, 101, 4096)):
        raw = rng.integers(300, 900, size=n).astype(np.int16)
        ncases.append(dict(key=f't{j}', lo=0, hi=n - 1))
        norm[f't{j}_raw'] = raw
        fixed = ref.Fast5.brute_remove(raw)
        norm[f't{j}_ref_brute'] = fixed
        with np.errstate(all='ignore'):
            norm[f't{j}_ref_norm_brute'] = ref.normalize_signal_mad(fixed)
            norm[f't{j}_ref_norm_none'] = ref.normalize_signal_mad(raw)
    np.savez_compressed(os.path.join(OUT, 'normalize.npz'), cases=json.dumps(ncases), **norm)

    # ---- expected signal (Squiggler._generate_signal) ---------------------------------------------
    seq = synth.random_flank(rng, 300)
    sq = ref.Squiggler('/nonexistent')
    np.savez_compressed(os.path.join(OUT, 'squiggle.npz'), seq=np.array(seq),
                        ref_signal=sq._generate_signal(seq))
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
