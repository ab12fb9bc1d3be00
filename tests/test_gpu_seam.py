"""GPU: the single-read seam with the reference's constructor, ``WarpSTR(flank_length, states,
endstate, repeat_mask, out_warp_path, reverse, read_name).run / .warp`` (caller.py:107-193), as
the reference's wrapper uses it: one object per read, built from the locus's automaton."""
import gc

import numpy as np
import pytest

from oracle import caller_oracle as co
from warpstr_b200 import synth
from warpstr_b200.automata import StateAutomata

pytestmark = pytest.mark.gpu


class _State:
    def __init__(self, kmer, value, seq_idx, idx):
        self.kmer, self.value, self.seq_idx, self.idx, self.incoming = kmer, value, seq_idx, idx, []


def _states_of(sta):
    """State objects shaped like the reference's (automata.py:10-18) from our flat automaton."""
    states = [_State(sta.kmers[i], float(sta.values[i]), int(sta.seq_idx[i]), i) for i in range(sta.n_states)]
    for i, s in enumerate(states):
        s.incoming = [states[int(p)] for p in sta.incoming_of(i)]
    return states


def test_seam_objects_of_successive_loci_do_not_share_tables(built_lib, oracle_c):
    """Two loci one after the other, the first locus's automaton collected in between: every read is
    called against its own locus's tables (the upload cache is keyed on content, not on object ids)."""
    from warpstr_b200.caller import CallerEngine, WarpSTR
    eng = CallerEngine()
    for name in ('HD', 'FMR1', 'AAAT', 'HD'):
        locus = synth.make_locus(name, seed=61)
        reads = synth.make_reads(locus, 3, seed=62)
        for r in reads:
            sta = StateAutomata(locus.reverse_regex if r.reverse else locus.template_regex)
            states = _states_of(sta)
            w = WarpSTR(110, states, sta.endstate, list(sta.mask), None, r.reverse, r.name, engine=eng)
            got = w.run(r.signal)
            want = co.run_read(r.signal, co.tables_from(sta), 110, r.reverse, impl='c')
            assert (got.seq, got.resc_seq, got.cost, got.resc_cost) == \
                (want.seq, want.resc_seq, want.cost, want.resc_cost), (name, r.name)
            t = w.warp(r.signal).trace
            assert np.array_equal(t, oracle_c.warp(r.signal, co.tables_from(sta),
                                                   np.zeros(len(r.signal), dtype=bool), 4, 110))
            del w, states, sta
            gc.collect()
    # same content -> same upload: the four HD/FMR1/AAAT strands, not one per read
    assert len(eng._seam_cache) <= 6


def test_call_arrays_results_do_not_alias_between_calls(built_lib):
    from warpstr_b200.caller import CallerEngine, pack_signals
    eng = CallerEngine()
    locus = synth.make_locus('HD', seed=5)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    ids = [eng.add_automaton(s, 110) for s in stas]
    outs = []
    for seed in (1, 2):
        reads = synth.make_reads(locus, 8, seed=seed)
        sigs = [r.signal for r in reads]
        if seed == 2:
            sigs[3] = sigs[3][:3]                      # too short: status 1, len -1, cost NaN
        host, off, lengths = pack_signals(sigs)
        aut = np.array([ids[int(r.reverse)] for r in reads], dtype=np.int32)
        rev = np.array([r.reverse for r in reads], dtype=np.uint8)
        outs.append(eng.call_arrays(host, off, lengths, aut, rev))
    first = {k: v.copy() for k, v in outs[0].items() if k in ('len1', 'len2', 'cost1', 'cost2', 'status')}
    reads = synth.make_reads(locus, 8, seed=1)
    host, off, lengths = pack_signals([r.signal for r in reads])
    again = eng.call_arrays(host, off, lengths, np.array([ids[int(r.reverse)] for r in reads], dtype=np.int32),
                            np.array([r.reverse for r in reads], dtype=np.uint8))
    for k, v in first.items():
        assert np.array_equal(outs[0][k], v) and np.array_equal(again[k], v, equal_nan=True)
    assert outs[1]['status'][3] == 1 and outs[1]['len2'][3] == -1 and np.isnan(outs[1]['cost2'][3])


def test_two_devices_in_one_process(built_lib, oracle_c):
    """One engine per GPU in the same process, calls interleaved: the staging slots and the kernels' launch
    attributes are per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from warpstr_b200.caller import CallerEngine
    locus = synth.make_locus('HD', seed=81)
    stas = [StateAutomata(locus.template_regex), StateAutomata(locus.reverse_regex)]
    reads = synth.make_reads(locus, 12, seed=82)
    want = [co.run_read(r.signal, co.tables_from(stas[int(r.reverse)]), 110, r.reverse, impl='c') for r in reads]
    engines = [CallerEngine(device=f'cuda:{d}') for d in (0, 1)]
    ids = [[e.add_automaton(s, 110) for s in stas] for e in engines]
    for rounds in range(3):
        for e, idd in zip(engines, ids):
            res = e.call_batch([r.signal for r in reads], [idd[int(r.reverse)] for r in reads], [r.reverse for r in reads])
            for g, w in zip(res, want):
                assert (g.seq, g.resc_seq, g.cost, g.resc_cost) == (w.seq, w.resc_seq, w.cost, w.resc_cost)
