"""CPU: the from-scratch automaton builder reproduces the reference's StateAutomata state by
state (k-mer, value, seq_idx, ordered incoming list, repeat mask, end state)."""
import json
import os

import numpy as np
import pytest

from warpstr_b200.automata import StateAutomata, parse_regex
from warpstr_b200.templates import reverse_uniq_sequence, reverse_complement

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_builder_matches_reference_goldens():
    z = np.load(os.path.join(GOLD, 'automata.npz'))
    meta = json.loads(str(z['meta']))
    assert len(meta) >= 30
    for m in meta:
        k = m['key']
        sta = StateAutomata(m['sequence'])
        assert sta.kmers == list(z[f'{k}_ref_kmers']), m['pattern']
        assert np.array_equal(sta.values, z[f'{k}_ref_values']), m['pattern']
        assert np.array_equal(sta.seq_idx, z[f'{k}_ref_seq_idx']), m['pattern']
        assert np.array_equal(sta.in_ptr, z[f'{k}_ref_in_ptr']), m['pattern']
        assert np.array_equal(sta.in_idx, z[f'{k}_ref_in_idx']), m['pattern']
        assert np.array_equal(np.array(sta.mask), z[f'{k}_ref_mask']), m['pattern']
        assert sta.endstate == int(z[f'{k}_ref_endstate'])
        assert (sta.repstart, sta.repend) == (int(z[f'{k}_ref_repstart']), int(z[f'{k}_ref_repend']))


def test_object_view_is_consistent():
    sta = StateAutomata('ACGTACGTAC' + '((CAGG){CAGM})(CA)' + 'TTGACCATGA')
    for i, s in enumerate(sta.states):
        assert s.idx == i and s.kmer == sta.kmers[i] and s.value == sta.values[i]
        assert [p.idx for p in s.incoming] == list(sta.incoming_of(i))
    assert sta.n_edges == sum(len(s.incoming) for s in sta.states)
    assert chr(sta.last_base[5]) == sta.kmers[5][-1]


def test_state_zero_is_the_only_source_and_seq_idx_is_sorted():
    for pat in ('(AAAT)', '(CAN)', '((CGG){AGG})', '(AC{GT}(TA))'):
        sta = StateAutomata('ACGTTGCAAGTC' + pat + 'GGATCCATTGCA')
        deg = np.diff(sta.in_ptr)
        assert deg[0] == 0 and (deg[1:] > 0).all()
        assert (np.diff(sta.seq_idx) >= 0).all()
        assert sta.endstate == sta.n_states - 1


def test_parse_regex_shapes():
    bases, succ, a, b = parse_regex('AC(GT)A')
    assert ''.join(bases) == 'ACGTA' and (a, b) == (2, 4)
    assert succ[3] == [2, 4]            # loop back to G, then on to A
    bases, succ, _, _ = parse_regex('A{C}G')
    assert succ[0] == [1, 2]            # optional C can be skipped


def test_reverse_strand_regex():
    assert reverse_uniq_sequence('(AGC)AAC{M}') == '{K}GTT(GCT)'
    assert reverse_complement('AACGT') == 'ACGTT'
