"""Locus regex -> k-mer state automaton, emitted as the flat tables the CUDA
kernels consume.

Behavioural mirror of the reference's ``StateAutomata`` (caller/automata.py:36-226):
the same regex grammar (plain bases, IUPAC letters = parallel alternatives,
``( )`` = one-or-more loop, ``{ }`` = optional group), the same context-split
k-mer states with the same numbering, ordered ``incoming`` lists, ``seq_idx``,
repeat mask and end state.  Built from scratch around index arrays instead of
linked objects; the object view (``.states[i].kmer/.value/.seq_idx/.idx/.incoming``)
is kept because the reference's caller seam (caller/caller.py:107-115) takes it.

Tables (all numpy, C-contiguous), S states and E edges:
  values   f64[S]   normalised pore level of each state's k-mer
  seq_idx  i32[S]   position of the k-mer's last base in the expanded regex
  in_ptr   i32[S+1], in_idx i32[E]   CSR of incoming states, in the reference's order
  rep_mask u8[S]    state belongs to the repeat region (automata.py:206)
  last_base u8[S]   ASCII of the k-mer's last base (what the decoded sequence emits)
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .pore_model import PoreModel, get_pore_model
from .templates import DNA_DICT


@dataclass
class State:
    """Object view of one k-mer state (reference: automata.py:10-18)."""
    kmer: str
    value: float
    seq_idx: int
    idx: int
    incoming: List['State'] = field(default_factory=list, repr=False)
    nextpos: List[int] = field(default_factory=list)


def parse_regex(sequence: str):
    """One pass over the locus regex producing single-base positions and their
    successor lists (reference: automata.py:57-150).

    Returns (bases, succ, repstart, repend): bases[p] is the base at position p,
    succ[p] the ordered list of positions that may follow it; repstart/repend
    delimit the repeat region (first ``(`` / last ``)``)."""
    bases = [sequence[0]]
    succ: List[List[int]] = [[]]
    loops = []           # open '(' : position (or positions, for an IUPAC opener) to loop back to
    optionals = []       # open '{' : position just before the optional group
    skip_from = []       # closed '{ }' groups waiting for the next plain base
    tails = [0]          # positions the next base chains from
    repstart = repend = -1

    for ch in sequence[1:]:
        here = len(bases)
        if ch == '(':
            loops.append(here)
            if repstart == -1:
                repstart = here
        elif ch == ')':
            repend = here
            back = loops.pop()
            targets = back if isinstance(back, list) else [back]
            for t in tails:
                succ[t].extend(targets)
        elif ch == '{':
            optionals.append(here - 1)
        elif ch == '}':
            skip_from.append(optionals.pop())
        elif ch in DNA_DICT:
            opens_group = bool((loops and here == loops[-1]) or (optionals and here == optionals[-1]))
            fresh = []
            for alt in DNA_DICT[ch]:
                pos = len(bases)
                bases.append(alt)
                succ.append([])
                fresh.append(pos)
                for t in tails:
                    succ[t].append(pos)
            if opens_group:
                # the loop has to return to every alternative of its first letter
                loops.pop()
                loops.append(list(fresh))
            tails = fresh
        else:
            bases.append(ch)
            succ.append([])
            for t in tails:
                succ[t].append(here)
            tails = [here]
            for s in skip_from:
                succ[s].append(here)
            skip_from = []
    return bases, succ, repstart, repend


class StateAutomata:
    """K-mer state automaton of one strand of one locus
    (reference: caller/automata.py:36-48)."""

    def __init__(self, sequence: str, pore_model: Optional[PoreModel] = None):
        pm = pore_model if pore_model is not None else get_pore_model()
        self.sequence = sequence
        self.kmersize = pm.kmersize
        bases, succ, self.repstart, self.repend = parse_regex(sequence)
        self._expand(bases, succ, pm)

    # -- context-split k-mer states (reference: automata.py:152-226) ---------------
    def _expand(self, bases, succ, pm: PoreModel):
        k = self.kmersize
        npos = len(bases)
        at_pos: List[List[int]] = [[] for _ in range(npos)]
        kmer: List[str] = []
        where: List[int] = []       # position of the k-mer's last base
        origin: List[int] = []      # position of the state this one was first reached from
        follow: List[List[int]] = []

        def spawn(word, pos, came_from):
            sid = len(kmer)
            kmer.append(word)
            where.append(pos)
            origin.append(came_from)
            follow.append([])
            at_pos[pos].append(sid)
            return sid

        # the first k-mer is read straight off the first k positions
        first = ''.join(bases[:k])
        todo = [spawn(first, k - 1, -1)]
        while todo:
            cur = todo.pop()
            stem = kmer[cur][1:]
            cur_pos = where[cur]
            for nxt in succ[cur_pos]:
                word = stem + bases[nxt]
                merged = False
                for other in at_pos[nxt]:
                    # merge only with a state spelling the same k-mer that was reached
                    # from the same position (keeps loop copies apart)
                    if kmer[other] == word and origin[other] == cur_pos:
                        merged = True
                        follow[cur].append(other)
                if not merged:
                    sid = spawn(word, nxt, cur_pos)
                    follow[cur].append(sid)
                    todo.append(sid)

        # flatten by position, then by creation order within a position
        order = [sid for pos in range(npos) for sid in at_pos[pos]]
        rank = {sid: i for i, sid in enumerate(order)}
        S = len(order)
        lo, hi = self.repstart - 1, self.repend + 10
        self.mask = [lo <= where[sid] <= hi for sid in order]
        self.endstate = rank[at_pos[-1][-1]] if at_pos[-1] else -1

        self.kmers = [kmer[sid] for sid in order]
        self.values = pm.get_values(self.kmers)
        self.seq_idx = np.array([where[sid] for sid in order], dtype=np.int32)
        nextpos = [[rank[t] for t in follow[sid]] for sid in order]
        incoming: List[List[int]] = [[] for _ in range(S)]
        for i in range(S):
            for t in nextpos[i]:
                incoming[t].append(i)
        self.in_ptr = np.zeros(S + 1, dtype=np.int32)
        self.in_ptr[1:] = np.cumsum([len(x) for x in incoming])
        self.in_idx = np.array([p for lst in incoming for p in lst], dtype=np.int32)
        self.rep_mask = np.array(self.mask, dtype=np.uint8)
        self.last_base = np.frombuffer(''.join(w[-1] for w in self.kmers).encode('ascii'),
                                       dtype=np.uint8).copy()

        self.states = [State(kmer=self.kmers[i], value=float(self.values[i]),
                             seq_idx=int(self.seq_idx[i]), idx=i, nextpos=nextpos[i])
                       for i in range(S)]
        for i, lst in enumerate(incoming):
            self.states[i].incoming = [self.states[p] for p in lst]

    @property
    def n_states(self) -> int:
        return len(self.kmers)

    @property
    def n_edges(self) -> int:
        return int(self.in_idx.shape[0])

    def incoming_of(self, j: int) -> np.ndarray:
        return self.in_idx[self.in_ptr[j]:self.in_ptr[j + 1]]
