"""TEST INFRASTRUCTURE ONLY -- the oracle over thousands of reads: ``run_read(impl='c', bulk=True)``
spread over a process pool on the host cores (the DP in C, the mid-stage on array slices).  Used by
the bulk parity tests and by bench.py's `parity` object; never by the product."""
import multiprocessing as mp
import os
from typing import List, Sequence, Tuple

import numpy as np

_CTX = {}


def _init(regexes: Sequence[str], flank_length: int, knobs):
    from oracle import caller_oracle as co
    from oracle import cdp
    from warpstr_b200.automata import StateAutomata
    cdp.build()
    _CTX['co'] = co
    _CTX['tb'] = [co.tables_from(StateAutomata(rx)) for rx in regexes]
    _CTX['F'] = flank_length
    _CTX['kn'] = co.Knobs(*knobs) if knobs is not None else None


def _one(job) -> Tuple:
    sig, aut, rev, want_traces = job
    co = _CTX['co']
    try:
        r = co.run_read(sig, _CTX['tb'][aut], _CTX['F'], bool(rev), _CTX['kn'], impl='c', bulk=True)
    except Exception as exc:                       # what the reference would raise for this read
        return ('error', type(exc).__name__)
    out = (r.seq, r.resc_seq, float(r.cost), float(r.resc_cost))
    if want_traces:
        out += (r.trace1.astype(np.int32), r.trace2.astype(np.int32))
    return out


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reads(regexes: Sequence[str], flank_length: int, signals: Sequence[np.ndarray], aut: Sequence[int],
              rev: Sequence[bool], knobs=None, want_traces: bool = False, workers: int = 0) -> List[Tuple]:
    """Oracle results, in order: (seq, resc_seq, cost, resc_cost[, trace1, trace2]) or ('error', type name).
    ``regexes[a]`` is the automaton regex of automaton id ``a``; ``knobs`` the Knobs fields as a tuple."""
    workers = workers or host_cores()
    jobs = [(np.ascontiguousarray(s, dtype=np.float64), int(a), bool(r), want_traces)
            for s, a, r in zip(signals, aut, rev)]
    if workers <= 1 or len(jobs) < 4:
        _init(regexes, flank_length, knobs)
        return [_one(j) for j in jobs]
    with mp.get_context('fork').Pool(workers, initializer=_init, initargs=(list(regexes), flank_length, knobs)) as pool:
        return pool.map(_one, jobs, chunksize=max(1, len(jobs) // (workers * 8)))
