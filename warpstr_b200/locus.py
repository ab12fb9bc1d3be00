"""Locus front-end: reference-genome slices without pysam, ``motif`` -> automaton regex.

Mirrors the reference's ``Locus.prepare_sequence`` (schemas/locus.py:48-100) and
``get_flanks`` / ``get_ref_pattern`` (squiggler/dna_sequence.py:31-56), which shell out to
``pysam.faidx``.  Here a FASTA + ``.fai`` index is read directly (plain, uncompressed FASTA).
Host-side, once per locus; no kernel involved.
"""
import os
from typing import Dict, Optional, Tuple

from .templates import reverse_complement


class FastaIndex:
    """samtools-faidx style random access: ``name length offset linebases linewidth`` per line."""

    def __init__(self, fasta_path: str):
        self.path = fasta_path
        fai = fasta_path + '.fai'
        if not os.path.exists(fasta_path):
            raise FileNotFoundError(fasta_path)
        self.index: Dict[str, Tuple[int, int, int, int]] = {}
        if os.path.exists(fai):
            with open(fai) as fh:
                for line in fh:
                    p = line.rstrip('\n').split('\t')
                    if len(p) >= 5:
                        self.index[p[0]] = (int(p[1]), int(p[2]), int(p[3]), int(p[4]))
        else:
            self._build()

    def _build(self):
        name, length, offset, lb, lw = None, 0, 0, 0, 0
        pos = 0
        with open(self.path, 'rb') as fh:
            for raw in fh:
                if raw.startswith(b'>'):
                    if name is not None:
                        self.index[name] = (length, offset, lb, lw)
                    name = raw[1:].split()[0].decode()
                    length, offset, lb, lw = 0, pos + len(raw), 0, 0
                else:
                    stripped = raw.rstrip(b'\r\n')
                    if lb == 0:
                        lb, lw = len(stripped), len(raw)
                    length += len(stripped)
                pos += len(raw)
            if name is not None:
                self.index[name] = (length, offset, lb, lw)

    def fetch(self, chrom: str, start: int, end: int) -> str:
        """1-based inclusive coordinates, like ``samtools faidx chrom:start-end``."""
        if chrom not in self.index:
            raise KeyError(f'sequence {chrom!r} not in {self.path}')
        length, offset, lb, lw = self.index[chrom]
        start = max(start, 1)
        end = min(end, length)
        if end < start:
            return ''
        a, b = start - 1, end
        first = offset + (a // lb) * lw + a % lb
        last = offset + ((b - 1) // lb) * lw + (b - 1) % lb + 1
        with open(self.path, 'rb') as fh:
            fh.seek(first)
            chunk = fh.read(last - first)
        return chunk.replace(b'\n', b'').replace(b'\r', b'').decode().upper()


def process_coord(coord: str) -> Tuple[str, int, int]:
    """'chr4:3,074,878-3,074,967' -> ('chr4', 3074878, 3074967)  (dna_sequence.py:16-23)."""
    chrom = coord.split(':')[0]
    start = int(coord.split(':')[1].split('-')[0].replace(',', ''))
    end = int(coord.split('-')[1].replace(',', ''))
    return chrom, start, end


def get_flanks(coord: str, ref: FastaIndex, flank_length: int, reverse: bool) -> Tuple[str, str]:
    chrom, start, end = process_coord(coord)
    left = ref.fetch(chrom, start - flank_length, start - 1)
    right = ref.fetch(chrom, end + 1, end + flank_length)
    if reverse:
        left, right = reverse_complement(right), reverse_complement(left)
    return left, right


def get_ref_pattern(coord: str, ref: FastaIndex) -> Tuple[str, str]:
    chrom, start, end = process_coord(coord)
    tmp = ref.fetch(chrom, start, end)
    return tmp, reverse_complement(tmp)


def _collapse(seq: str, motif: str, annotate: bool) -> str:
    """Replace runs of >= 2 consecutive copies of ``motif`` by ``(motif)`` (or ``(motif)[n]``),
    single copies stay literal -- the reference's split/join walk (locus.py:59-96)."""
    out = ''
    run = 0
    for ch in '-'.join(seq.split(motif)):
        if ch == '-':
            run += 1
            continue
        if run > 1:
            out += f'({motif})[{run}]' if annotate else f'({motif})'
        elif run == 1:
            out += motif
        out += ch
        run = 0
    if run > 1:
        out += f'({motif})[{run}]' if annotate else f'({motif})'
    elif run == 1:
        out += motif
    return out


def prepare_sequence(ref_seq: str, motif: str, name: str = '', coord: str = '') -> Tuple[str, str]:
    """Reference repeat region + comma separated motifs -> (automaton regex, annotated note)."""
    seq = ref_seq.upper()
    noting = seq
    if len(seq) <= 1:
        raise ValueError(f'Reference repeat sequence is empty for {name} - check coord {coord}')
    if not motif:
        raise ValueError(f'Motif is not defined for {name}. Define either motif or sequence')
    for m in motif.split(','):
        seq = _collapse(seq, m, False)
        noting = _collapse(noting, m, True)
    return seq, noting


def locus_sequence(coord: str, motif: Optional[str], sequence: Optional[str], reference_path: str,
                   name: str = '') -> Tuple[str, Optional[str]]:
    """What ``Locus.__init__`` settles on (locus.py:39-46): the configured regex, or one derived
    from the motifs and the reference genome."""
    if sequence:
        return sequence.upper(), None
    ref = FastaIndex(reference_path)
    seq, noting = prepare_sequence(get_ref_pattern(coord, ref)[0], motif or '', name, coord)
    return seq.upper(), noting


def write_expected_signals(locus_path: str, coord: str, reference_path: str, flank_length: int,
                           pore_model=None, device: str = 'cuda') -> Dict[str, str]:
    """The expected-signal step of one locus (reference: Squiggler.process_locus,
    squiggler/Squiggler.py:30-50, 68-75): ``expected_signals/sequences.csv`` with the four flanks
    and the reference pattern of both strands, plus one ``<name>.txt`` of pore-model levels per
    sequence (levels looked up on the GPU).  Returns the sequences by name."""
    import numpy as np
    from . import templates as tmpl
    from .pore_model import get_pore_model
    pm = pore_model or get_pore_model()
    ref = FastaIndex(reference_path)
    lf_t, rf_t = get_flanks(coord, ref, flank_length, reverse=False)
    lf_r, rf_r = get_flanks(coord, ref, flank_length, reverse=True)
    tmpp, revp = get_ref_pattern(coord, ref)
    seqs = dict(zip(tmpl.LOCUS_NAMES, (lf_t, rf_t, lf_r, rf_r, tmpp, revp)))
    out_dir = os.path.join(locus_path, tmpl.LOCUS_INFO_SUBDIR)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, tmpl.LOCUS_FLANKS), 'w') as fh:
        fh.write('type,sequence\n')
        for name in tmpl.LOCUS_NAMES:
            fh.write(name + ',' + seqs[name].upper() + '\n')
    for name in tmpl.LOCUS_NAMES:
        np.savetxt(os.path.join(out_dir, name + '.txt'), pm.generate_signal(seqs[name], device=device), fmt='%f')
    return seqs
