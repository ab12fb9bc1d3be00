"""The caller seam: ``CallerWrapper(locus, threads).run(workload) -> List[CallerResult]``.

Drop-in for the reference's ``CallerWrapper`` (caller/wrapper.py:57-248): builds the
template- and reverse-strand automata from the locus regex and its flanks, runs every read
of the workload (order preserved) and offers the same helpers for complex loci
(``break_into_units``, ``collapse_repeats``).  Where the reference fans reads out to a
``multiprocessing.Pool`` (:104-120), this sends the whole workload to the GPU in one batch;
``threads`` is accepted for signature compatibility and ignored.
"""
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import templates as tmpl
from .automata import StateAutomata
from .caller import CallerEngine, CallerResult
from .config import CallerConfig, Config, RescalerConfig
from .pore_model import PoreModel, get_pore_model


@dataclass
class ReadSignal:
    """schemas/readsignal.py:6-10"""
    name: str
    reverse: bool
    signal: np.ndarray


@dataclass
class Flank:
    left: str
    right: str


@dataclass
class Flanks:
    template: Flank
    reverse: Flank


def load_flanks(path: str) -> Flanks:
    """Flanks written by the expected-signal step (``expected_signals/sequences.csv``,
    lines ``type,sequence``; reference: extractor/tr_extractor.py:108-140)."""
    fname = os.path.join(path, tmpl.LOCUS_INFO_SUBDIR, tmpl.LOCUS_FLANKS)
    seqs = {}
    with open(fname, 'r') as fh:
        next(fh)
        for line in fh:
            parts = line.strip().split(',')
            if len(parts) == 2:
                seqs[parts[0]] = parts[1]
    try:
        return Flanks(template=Flank(seqs['left_flank_template'], seqs['right_flank_template']),
                      reverse=Flank(seqs['left_flank_reverse'], seqs['right_flank_reverse']))
    except KeyError as exc:
        raise KeyError(f'{fname} lacks the flank {exc}')


def flanks_from_template(left: str, right: str) -> Flanks:
    """Reverse-strand flanks are the swapped reverse complements (dna_sequence.py:54-55)."""
    return Flanks(template=Flank(left, right),
                  reverse=Flank(tmpl.reverse_complement(right), tmpl.reverse_complement(left)))


@dataclass
class Locus:
    """The fields of the reference's Locus that the caller reads (schemas/locus.py:8-46).
    ``sequence`` must be given; deriving it from ``motif`` needs the reference genome and is
    outside this path."""
    name: str
    sequence: str
    flank_length: int = 110
    path: str = ''
    coord: str = ''

    def __post_init__(self):
        self.sequence = self.sequence.upper()


class CallerWrapper:
    def __init__(self, locus: Locus, threads: int = 1, flanks: Optional[Flanks] = None,
                 caller_config: Optional[CallerConfig] = None, rescaler_config: Optional[RescalerConfig] = None,
                 pore_model: Optional[PoreModel] = None, engine: Optional[CallerEngine] = None,
                 config: Optional[Config] = None):
        if config is not None:
            caller_config = caller_config or config.caller_config
            rescaler_config = rescaler_config or config.rescaler_config
        self.locus = locus
        self.threads = threads
        self.caller_config = caller_config or CallerConfig()
        self.pore_model = pore_model or get_pore_model()
        self.flanks = flanks if flanks is not None else load_flanks(locus.path)
        template_seq, reverse_seq = self.get_seqs(locus.sequence, self.flanks)
        if locus.path and os.path.isdir(os.path.join(locus.path, tmpl.SUMMARY_SUBDIR)):
            self.check_high_similarity(locus.sequence)
        self.units, self.repeat_units, self.offsets = self.break_into_units(locus.sequence)
        self.temp_sta = StateAutomata(template_seq, self.pore_model)
        self.rev_sta = StateAutomata(reverse_seq, self.pore_model)
        self.engine = engine or CallerEngine(self.caller_config, rescaler_config)
        self._temp_id = self.engine.add_automaton(self.temp_sta, locus.flank_length)
        self._rev_id = self.engine.add_automaton(self.rev_sta, locus.flank_length)

    # -- sequences (wrapper.py:72-84) --------------------------------------------------------------
    def get_seqs(self, sequence: str, flanks: Flanks):
        tmp = flanks.template.left + sequence + flanks.template.right
        rev = flanks.reverse.left + self.reverse_uniq_sequence(sequence) + flanks.reverse.right
        return tmp, rev

    @staticmethod
    def reverse_uniq_sequence(sequence: str) -> str:
        return tmpl.reverse_uniq_sequence(sequence)

    # -- the call (wrapper.py:104-120) ---------------------------------------------------------------
    def run(self, workload: Sequence[ReadSignal]) -> List[CallerResult]:
        signals = [np.ascontiguousarray(r.signal, dtype=np.float64) for r in workload]
        reverse = [bool(r.reverse) for r in workload]
        aut = [self._rev_id if rv else self._temp_id for rv in reverse]
        return self.engine.call_batch(signals, aut, reverse)

    def run_raw(self, raws: Sequence[np.ndarray], windows, reverse: Sequence[bool],
                spike_removal: Optional[str] = None) -> List[CallerResult]:
        """``get_workload`` + ``run`` for reads still in DAC counts (wrapper.py:44-54, 104-120): normalised
        and called on the device in one chain, order preserved."""
        reverse = [bool(r) for r in reverse]
        aut = [self._rev_id if rv else self._temp_id for rv in reverse]
        return self.engine.call_raw_batch(raws, windows, aut, reverse,
                                          spike_removal or self.caller_config.spike_removal)

    # -- state similarity report (wrapper.py:122-160) ------------------------------------------------
    def check_high_similarity(self, sequence: str):
        diffs = {'template': self.pore_model.get_diffs_for_all(sequence),
                 'reverse': self.pore_model.get_diffs_for_all(self.reverse_uniq_sequence(sequence))}
        out_path = os.path.join(self.locus.path, tmpl.SUMMARY_SUBDIR, 'state_similarity.csv')
        try:
            with open(out_path, 'w') as fh:
                fh.write('pattern,strand,mean_diff,median_diff\n')
                for strand, table in diffs.items():
                    for pat, (mean_d, med_d) in table.items():
                        fh.write(f'{pat},{strand},{mean_d:.3f},{med_d:.3f}\n')
        except OSError:
            pass
        limit = self.caller_config.min_state_similarity
        problems = {strand: [dict(pattern=pat, mean_diff=v[0], median_diff=v[1])
                             for pat, v in table.items() if limit > v[0] or limit > v[1]]
                    for strand, table in diffs.items()}
        for p in problems['template']:
            print('Warning: Template has repeat unit {} with high state similarity'.format(p['pattern']))
        for p in problems['reverse']:
            print('Warning: high similarity of state values in reverse pattern {}'.format(p['pattern']))
        return problems['template'], problems['reverse']

    # -- complex loci (wrapper.py:162-248) -----------------------------------------------------------
    def break_into_units(self, template: str):
        """Top-level bracketed groups of the locus regex, the plain-base strings each can emit
        per loop iteration, and the number of unbracketed bases in front of each group."""
        opened: List[int] = []
        units: List[str] = []
        offsets: List[int] = []
        plain = 0
        for pos, ch in enumerate(template):
            if ch in '({':
                opened.append(pos)
            elif ch in ')}':
                begin = opened.pop()
                if not opened:
                    units.append(template[begin:pos + 1])
                    offsets.append(plain)
                    plain = 0
            elif not opened:
                plain += 1

        repeat_units: List[List[str]] = []
        for unit in units:
            variants: List[str] = []
            cur = ''
            for ch in unit:
                if ch in '()':
                    continue
                if ch == '{':
                    variants.append(cur)
                    cur = ''
                elif ch == '}':
                    variants.extend([v + cur for v in list(variants)])
                    cur = ''
                else:
                    cur += ch
            if cur:
                variants.append(cur)
            expanded: List[str] = []
            for var in variants:
                outs = ['']
                for ch in var:
                    if ch in tmpl.DNA_DICT:
                        outs = [o + alt for alt in tmpl.DNA_DICT[ch] for o in outs]
                    else:
                        outs = [o + ch for o in outs]
                expanded.extend(outs)
            repeat_units.append(expanded)
        return units, repeat_units, offsets

    def collapse_repeats(self, seq: str):
        """Counts of every repeat-unit variant in a called sequence (wrapper.py:220-248)."""
        results = [[0] * len(u) for u in self.repeat_units]
        rest = seq
        for n, (variants, off) in enumerate(zip(self.repeat_units, self.offsets)):
            rest = rest[off:]
            while rest:
                nxt = None
                for v, cand in enumerate(variants):
                    if cand == rest[:len(cand)]:
                        results[n][v] += 1
                        nxt = rest[len(cand):]      # the last matching variant decides what is consumed
                if nxt is None:
                    break
                rest = nxt
        return results


def get_raw_workload(df_overview, path: str):
    """The saved reads of a locus as they are on disk: (names, reverse flags, raw int16 signals, windows).
    The raw signal of every saved read comes from its annotated single-read fast5
    (``<locus>/fast5/<run_id>/annot/<read>.fast5``, what the reference's extraction step
    writes) or, for a caller-only run prepared straight from a sequencing run's multi-read
    files, from the file named in the row's ``fast5_path`` column (see caller_only.py)."""
    from .fast5 import raw_signal
    names, revs, raws, wins = [], [], [], []
    has_src = 'fast5_path' in df_overview.columns
    for row in df_overview.itertuples():
        if row.saved:
            fpath = os.path.join(path, tmpl.FAST5_SUBDIR, str(row.run_id), tmpl.ANNOT_SUBDIR, row.Index + '.fast5')
            if os.path.exists(fpath):
                raws.append(raw_signal(fpath))
            elif has_src and isinstance(row.fast5_path, str) and os.path.exists(row.fast5_path):
                raws.append(raw_signal(row.fast5_path, row.Index))
            else:
                raise FileNotFoundError(f'no fast5 for read {row.Index}: {fpath}')
            names.append(row.Index)
            revs.append(bool(row.reverse))
            wins.append((int(row.l_start_raw), int(row.r_end_raw)))
    return names, revs, raws, wins


def get_workload(df_overview, path: str, spike_removal: str = 'Brute') -> List[ReadSignal]:
    """Reads of the locus with their normalised STR windows (reference: wrapper.py:44-54): all reads are
    spike-filtered, MAD normalised and sliced in one GPU launch and returned as host float64 arrays, which
    is what the reference's ``ReadSignal`` holds.  ``main_wrapper`` does not take this detour: it hands the
    raw reads to ``CallerWrapper.run_raw``, which keeps the windows on the device."""
    from .normalize import normalize_windows
    names, revs, raws, wins = get_raw_workload(df_overview, path)
    sigs = normalize_windows(raws, wins, spike_removal)
    return [ReadSignal(n, r, s) for n, r, s in zip(names, revs, sigs)]


def main_wrapper(locus: Locus, threads: int = 1, config: Optional[Config] = None,
                 workload: Optional[Sequence[ReadSignal]] = None, flanks: Optional[Flanks] = None,
                 engine: Optional[CallerEngine] = None):
    """The caller step of one locus (reference: wrapper.py:17-41): overview.csv in, calls on the
    GPU, overview.csv / FASTA / complex-unit CSV out.  Plots are not produced.  ``workload``
    (one ReadSignal per saved overview row, in row order) bypasses the fast5 reading."""
    from .overview import load_overview, store_collapsed, store_results
    overview_path, df_overview = load_overview(locus.path)
    cc = config.caller_config if config is not None else CallerConfig()
    call_wrapper = CallerWrapper(locus, threads, flanks=flanks, config=config, engine=engine)
    if workload is None:
        # raw int16 reads -> normalisation kernel -> caller, all on the device
        names, revs, raws, wins = get_raw_workload(df_overview, locus.path)
        results = call_wrapper.run_raw(raws, wins, revs, cc.spike_removal)
        workload = [ReadSignal(n, r, None) for n, r in zip(names, revs)]
    else:
        results = call_wrapper.run(workload)
    seq_results = [(r.seq, r.resc_seq) for r in results]
    cost_results = [(r.cost, r.resc_cost) for r in results]
    df_overview = store_results(overview_path, df_overview, seq_results, cost_results, locus.path)
    df_collapsed = None
    if len(call_wrapper.units) > 1:
        print(f'Running complex genotyping as complex repeat units present: {call_wrapper.units}')
        collapsed = [call_wrapper.collapse_repeats(s[1]) for s in seq_results]
        df_collapsed = store_collapsed(collapsed, call_wrapper.units, call_wrapper.repeat_units,
                                       [bool(r.reverse) for r in workload], locus.path)
    return df_overview, df_collapsed
