"""Host evaluation of the stage between the two DP passes, for the reads the device flags.

What the reference does between ``warp(signal)`` and ``warp(rescaled_signal, badmask)``
(caller/caller.py:123-125 and 132-139): per-run statistics of the alignment
(:65-96), the smoothing-spline rescale (:304-318) and the t-test segmentation that
marks badly resolved repeat windows (:330-421).  Vectorised per read; the numpy
reductions that decide bit-level results (pairwise ``mean``/``std`` per run,
``splrep``/``splev``) are the same calls the reference makes, and the t-test squares with
the host libm's ``pow`` like the reference's ``np.float64 ** 2``.

The device runs this stage for every configuration (``wstr_call_batch``).  This module is only
reached for a read the device has flagged -- a smoothing spline that needs interior knots
(``WSTR_READ_SPLINE_KNOTS``), a t-test decision within rounding distance of flipping
(``d_ttest_ties``), a status that the reference answers with an exception (to raise its type) --
and for ``engine='host'`` in the tests.  scipy is imported when such a read occurs, not before.
"""
import math
from dataclasses import dataclass

import numpy as np

from .config import CallerConfig, RescalerConfig


class ReadError(Exception):
    """A per-read failure with the exception type the reference would raise."""

    def __init__(self, kind, msg):
        super().__init__(msg)
        self.kind = kind


@dataclass
class Runs:
    """Run-length view of a trace (reference: WarpResult.state_transitions, caller.py:58-60)."""
    states: np.ndarray     # state index of each run
    starts: np.ndarray     # first sample of each run
    ends: np.ndarray       # one past the last sample

    @property
    def n(self) -> int:
        return int(self.states.shape[0])


def run_lengths(trace: np.ndarray) -> Runs:
    trace = np.asarray(trace)
    cut = np.flatnonzero(trace[1:] != trace[:-1]) + 1
    starts = np.concatenate(([0], cut))
    ends = np.concatenate((cut, [trace.shape[0]]))
    return Runs(states=trace[starts].astype(np.int64), starts=starts, ends=ends)


@dataclass
class Alignment:
    state_value: np.ndarray    # per run (or per distinct state if reps_as_one)
    expected: np.ndarray
    good: np.ndarray           # bool: usable for rescaling (caller.py:23-39)

    def cost(self, start: int, end: int) -> float:
        # caller.py:138-139: mean |state_value - expected| over alignment[start:end]
        return np.mean(np.abs(self.state_value[start:end] - self.expected[start:end]).tolist())


def create_alignment(trace: np.ndarray, runs: Runs, x: np.ndarray, values: np.ndarray,
                     cc: CallerConfig, rc: RescalerConfig) -> Alignment:
    if rc.method not in ('mean', 'median'):
        raise KeyError(f'Invalid alignment method: {rc.method}')
    collapse = np.mean if rc.method == 'mean' else np.median
    if rc.reps_as_one:
        ids = np.unique(runs.states)
        groups = [x[trace == s] for s in ids]
        expected = values[ids]
    else:
        groups = [x[a:b] for a, b in zip(runs.starts, runs.ends)]
        expected = values[runs.states]
    sv = np.array([collapse(g) for g in groups], dtype=np.float64)
    good = np.zeros(len(groups), dtype=bool)
    for n, g in enumerate(groups):
        if len(g) >= cc.min_values_per_state and np.std(g) < rc.max_std and \
                abs(expected[n] - sv[n]) <= rc.threshold:
            good[n] = True
    return Alignment(state_value=sv, expected=np.asarray(expected, dtype=np.float64), good=good)


def rescale_signal(x: np.ndarray, al: Alignment) -> np.ndarray:
    xs = al.state_value[al.good]
    ys = al.expected[al.good]
    order = np.argsort(xs, kind='stable')
    xs, ys = xs[order], ys[order]
    from scipy import interpolate        # only here: the product path proper never imports scipy
    try:
        tck = interpolate.splrep(xs.tolist(), ys.tolist(), s=len(xs))
    except Exception as exc:                      # FITPACK input errors (caller.py:311)
        raise ReadError(type(exc), str(exc))
    return np.asarray(interpolate.splev(x, tck))


def _tstats(x: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Sliding two-sample statistic of caller.py:347-354 at every centre c in [lo, hi]:
    windows x[c-3:c] and x[c:c+3], population std, sequential 3-term sums as numpy
    performs them for n < 8."""
    c = np.arange(lo, hi + 1)
    w = x[(c[:, None] - 3) + np.arange(6)[None, :]]
    a, b = w[:, :3], w[:, 3:]

    def stats(m):
        mean = ((m[:, 0] + m[:, 1]) + m[:, 2]) / 3.0
        d = m - mean[:, None]
        var = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]) / 3.0
        sd = np.sqrt(var)
        return mean, sd

    ma, sa = stats(a)
    mb, sb = stats(b)
    # the reference squares numpy scalars, `np.std(a) ** 2` (caller.py:351): that is libm's pow(x, 2.0),
    # which need not be x*x in the last bit -- use the host's pow so that a tie is decided as the
    # reference decides it on this machine
    sa2 = np.fromiter((math.pow(v, 2.0) for v in sa.tolist()), dtype=np.float64, count=len(sa))
    sb2 = np.fromiter((math.pow(v, 2.0) for v in sb.tolist()), dtype=np.float64, count=len(sb))
    sd = np.sqrt((sa2 + sb2) / 3)
    sd = np.where(sd == 0, sd + 0.0000001, sd)
    return (ma - mb) / sd


def _count_segments(t: np.ndarray) -> int:
    # caller.py:357-378 (the border positions themselves are not used, only their number)
    n_borders = 0
    rising = False
    prev = t[0]
    for v in t.tolist():
        if v > 3 or v < -3:
            if (v > 3 and v >= prev) or (v < -3 and v <= prev):
                rising = True
            else:
                if rising:
                    n_borders += 1
                rising = False
        elif rising:
            n_borders += 1
            rising = False
        prev = v
    return n_borders - 1


def find_event_borders(rep_mask: np.ndarray, runs: Runs, cc: CallerConfig):
    """caller.py:381-406 on the run-length view.  Returns (start, end, start_idx, end_idx,
    bounds): run indices of the repeat region and the sample indices of every
    ``states_in_segment``-th run boundary inside it."""
    sis = cc.states_in_segment
    in_rep = np.flatnonzero(rep_mask[runs.states])
    if in_rep.size == 0:
        raise ReadError(IndexError, 'list index out of range')       # trues[0]
    start, end = int(in_rep[0]), int(in_rep[-1])

    def span(a, b):
        if b >= runs.n:
            raise ReadError(IndexError, 'index out of bounds')       # state_transitions[end]
        ra = int(np.flatnonzero(runs.states == runs.states[a])[0])
        rb = int(np.flatnonzero(runs.states == runs.states[b])[-1])
        lo, hi = int(runs.starts[ra]), int(runs.ends[rb]) - 1
        cuts = runs.ends[ra:rb] - 1 if rb > ra else np.zeros(0, dtype=np.int64)
        return lo, hi, cuts

    lo, hi, cuts = span(start, end)
    extra = (len(cuts) - 1) % sis
    if extra > 0:
        end = end + (sis - extra)
        lo, hi, cuts = span(start, end)
    bounds = cuts[::sis]
    return start, end, lo, hi, bounds


def mask_bad_repeats(x: np.ndarray, rep_mask: np.ndarray, runs: Runs, cc: CallerConfig):
    """caller.py:330-344, 409-421.  Returns (start, end, badmask bool[T])."""
    start, end, lo, hi, bounds = find_event_borders(rep_mask, runs, cc)
    T = x.shape[0]
    if len(bounds) == 0:
        raise ReadError(IndexError, 'list index out of range')       # bounds[0]
    bad = np.zeros(T, dtype=bool)
    if len(bounds) > 1:
        b0, b1 = int(bounds[0]), int(bounds[-1])
        # a window is x[b_n-3 : b_{n+1}+3]; python clips the slice at T and wraps a negative
        # start, after which segment() indexes an empty list (caller.py:336,358-361)
        last_c = min(b1, T - 3)
        if b0 - 3 < 0 or last_c < b0:
            raise ReadError(IndexError, 'list index out of range')
        t = _tstats(x, b0, last_c)
        for n in range(len(bounds) - 1):
            seg = t[int(bounds[n]) - b0:int(bounds[n + 1]) - b0 + 1]
            if len(seg) == 0:
                raise ReadError(IndexError, 'list index out of range')
            if _count_segments(seg) >= cc.states_in_segment + 1:
                bad[int(bounds[n]):int(bounds[n + 1])] = True
    return start, end, bad


@dataclass
class PassResult:
    runs: Runs
    alignment: Alignment
    rescaled: np.ndarray
    start: int
    end: int
    badmask: np.ndarray
    cost: float


def after_pass(trace: np.ndarray, x: np.ndarray, values: np.ndarray, rep_mask: np.ndarray,
               cc: CallerConfig, rc: RescalerConfig, mask_on_rescaled: bool) -> PassResult:
    """Everything the reference computes from one pass's trace.  First pass
    (caller.py:123-126): the mask is derived from the input signal.  Second pass
    (:132-135): from the re-rescaled signal (only start/end are used)."""
    runs = run_lengths(trace)
    al = create_alignment(trace, runs, x, values, cc, rc)
    resc = rescale_signal(x, al)
    start, end, bad = mask_bad_repeats(resc if mask_on_rescaled else x, rep_mask, runs, cc)
    return PassResult(runs=runs, alignment=al, rescaled=resc, start=start, end=end, badmask=bad,
                      cost=al.cost(start, end))
