"""Deterministic synthetic STR reads (SURVEY.md section 8d is the spec; the reference
ships no simulator).

A locus is two random (seeded) plain-ACGT flanks around a regex; a read is one
realisation of the regex (allele), turned into pore-model levels of its sliding
6-mers, each level held for a random dwell, plus Gaussian noise, in the normalised
units the caller works in.  Reverse-strand reads realise the reverse-complemented
sequence, which is what the reverse automaton (wrapper.py:72-84) expects.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from .pore_model import PoreModel, get_pore_model
from .templates import reverse_complement, reverse_uniq_sequence


@dataclass
class SynthLocus:
    name: str
    sequence: str                 # automaton regex of the template strand
    left: str                     # template-strand flanks
    right: str
    units: Tuple                  # generator recipe, see ``draw_allele``
    flank_length: int = 110

    @property
    def template_regex(self) -> str:
        return self.left + self.sequence + self.right

    @property
    def reverse_regex(self) -> str:
        # flanks of the reverse strand are the swapped reverse complements
        # (dna_sequence.py:54-55), the regex is mirrored (wrapper.py:78-84)
        return reverse_complement(self.right) + reverse_uniq_sequence(self.sequence) + \
            reverse_complement(self.left)


def random_flank(rng: np.random.Generator, n: int) -> str:
    return ''.join(rng.choice(list('ACGT'), n))


# generator recipes: a list of parts; each part is either a literal string or
# (unit, lo, hi[, interruption, p]) = unit repeated U{lo..hi} times, each copy replaced
# by `interruption` with probability p.
RECIPES: Dict[str, Tuple[str, Tuple]] = {
    'AAAT': ('(AAAT)', (('AAAT', 9, 12),)),
    'HD': ('(AGC)AACAGCCGCCAC(CGC)', (('AGC', 30, 45), 'AACAGCCGCCAC', ('CGC', 7, 12))),
    'FMR1': ('((CGG){AGG})', (('CGG', 25, 35, 'CGGAGG', 0.06),)),
    'FMR1_MGG': ('(MGG)', (('CGG', 25, 35, 'AGG', 0.06),)),
    'DM2': ('((CAGG){CAGM})(CAGA)(CA)', (('CAGG', 8, 20, 'CAGGCAGA', 0.1), ('CAGA', 5, 12), ('CA', 10, 20))),
    'C9ORF72_100': ('(GGGGCC)', (('GGGGCC', 90, 110),)),
    'C9ORF72_300': ('(GGGGCC)', (('GGGGCC', 280, 320),)),
    'C9ORF72_1000': ('(GGGGCC)', (('GGGGCC', 950, 1000),)),
    'CAN': ('(CAN)', (('CAG', 15, 30, 'CAA', 0.2),)),
    'RFC1': ('(AARRG)', (('AAGGG', 10, 30, 'AAAAG', 0.3),)),
}


def make_locus(name: str, seed: int = 0, flank_length: int = 110,
               recipe: Optional[str] = None) -> SynthLocus:
    regex, units = RECIPES[recipe or name]
    rng = np.random.default_rng([seed, 0x10C05])
    return SynthLocus(name=name, sequence=regex, left=random_flank(rng, flank_length),
                      right=random_flank(rng, flank_length), units=units,
                      flank_length=flank_length)


def draw_allele(rng: np.random.Generator, units: Tuple) -> str:
    out = []
    for part in units:
        if isinstance(part, str):
            out.append(part)
            continue
        unit, lo, hi = part[0], part[1], part[2]
        n = int(rng.integers(lo, hi + 1))
        if len(part) > 3:
            alt, p = part[3], part[4]
            out.extend(alt if rng.random() < p else unit for _ in range(n))
        else:
            out.append(unit * n)
    return ''.join(out)


def squiggle(rng: np.random.Generator, bases: str, pm: PoreModel, noise: float = 0.15,
             dwell: Tuple[int, int] = (5, 13)) -> np.ndarray:
    """Levels of the sliding 6-mers of ``bases`` -> noisy, dwell-expanded signal."""
    codes = np.frombuffer(bases.encode('ascii'), dtype=np.uint8)
    lut = np.zeros(256, dtype=np.int64)
    for i, b in enumerate('ACGT'):
        lut[ord(b)] = i
    c = lut[codes]
    k = pm.kmersize
    idx = np.zeros(len(c) - k + 1, dtype=np.int64)
    for p in range(k):
        idx = idx * 4 + c[p:len(c) - k + 1 + p]
    levels = pm.table[idx]
    dw = rng.integers(dwell[0], dwell[1] + 1, size=len(levels))
    sig = np.repeat(levels, dw)
    return sig + rng.normal(0.0, noise, size=sig.shape[0])


@dataclass
class SynthRead:
    name: str
    reverse: bool
    signal: np.ndarray            # float64, normalised
    truth_len: int                # STR length in nucleotides of the drawn allele
    locus: int = 0


def make_reads(locus: SynthLocus, n: int, seed: int = 0, noise: float = 0.15,
               reverse_fraction: float = 0.5, pm: Optional[PoreModel] = None,
               locus_id: int = 0) -> List[SynthRead]:
    pm = pm or get_pore_model()
    rng = np.random.default_rng([seed, 0x5EAD])
    reads = []
    for i in range(n):
        allele = draw_allele(rng, locus.units)
        seq = locus.left + allele + locus.right
        rev = bool(rng.random() < reverse_fraction)
        if rev:
            seq = reverse_complement(seq)
        sig = squiggle(rng, seq, pm, noise)
        reads.append(SynthRead(name=f'{locus.name}_{seed}_{i}', reverse=rev, signal=sig,
                               truth_len=len(allele), locus=locus_id))
    return reads


def to_raw_int16(rng: np.random.Generator, norm: np.ndarray, pad: int = 8192,
                 spike_rate: float = 1e-4, pm: Optional[PoreModel] = None) -> Tuple[np.ndarray, int, int]:
    """Inverse of the normalisation for exercising the raw-signal kernel: map a
    normalised window to DAC counts, embed it in a longer read and inject spikes.
    Returns (raw int16[N], l_start_raw, r_end_raw)."""
    pm = pm or get_pore_model()
    left = pad // 2
    filler = pm.table[rng.integers(0, len(pm.table), size=pad // 9 + 2)]
    fill = np.repeat(filler, 9)[:pad] + rng.normal(0, 0.15, size=pad)
    full = np.concatenate([fill[:left], norm, fill[left:]])
    raw = np.rint((90.8717 + 9.8354 * full) * 5.85)
    spikes = rng.random(raw.shape[0]) < spike_rate
    raw[spikes] = np.where(rng.random(int(spikes.sum())) < 0.5, 1500, 100)
    return raw.astype(np.int16), left, left + len(norm) - 1


def make_read_batch(locus: SynthLocus, n: int, seed: int = 0, noise: float = 0.15,
                    reverse_fraction: float = 0.5, pm: Optional[PoreModel] = None,
                    dwell: Tuple[int, int] = (5, 13)):
    """Vectorised generator for large batches: same model as :func:`make_reads`, different
    random stream.  Returns (signal f64 concatenated with even starts, offsets i64[n],
    lengths i32[n], reverse u8[n], truth_len i32[n])."""
    pm = pm or get_pore_model()
    rng = np.random.default_rng([seed, 0xBA7C4])
    k = pm.kmersize
    seqs, rev, truth = [], np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.int32)
    rflags = rng.random(n) < reverse_fraction
    for i in range(n):
        allele = draw_allele(rng, locus.units)
        s = locus.left + allele + locus.right
        if rflags[i]:
            s = reverse_complement(s)
            rev[i] = 1
        truth[i] = len(allele)
        seqs.append(s)
    nb = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
    codes = np.frombuffer(''.join(seqs).encode('ascii'), dtype=np.uint8)
    lut = np.zeros(256, dtype=np.int64)
    for i, b in enumerate('ACGT'):
        lut[ord(b)] = i
    c = lut[codes]
    idx = np.zeros(len(c) - k + 1, dtype=np.int64)
    for p in range(k):
        idx = idx * 4 + c[p:len(c) - k + 1 + p]
    # drop the k-mers that straddle two reads
    starts = np.concatenate(([0], np.cumsum(nb)[:-1]))
    keep = np.ones(len(idx), dtype=bool)
    for p in range(1, k):
        bad = starts[1:] - p
        keep[bad[bad >= 0]] = False
    levels = pm.table[idx[keep]]
    nk = nb - k + 1                                   # k-mers per read
    dw = rng.integers(dwell[0], dwell[1] + 1, size=len(levels))
    kstart = np.concatenate(([0], np.cumsum(nk)))
    cum = np.concatenate(([0], np.cumsum(dw)))
    lengths = (cum[kstart[1:]] - cum[kstart[:-1]]).astype(np.int32)
    sig = np.repeat(levels, dw)
    sig += rng.normal(0.0, noise, size=sig.shape[0])
    # re-pack with even starts (the kernels want 16-byte aligned reads)
    padded = (lengths.astype(np.int64) + 1) & ~1
    offsets = np.zeros(n, dtype=np.int64)
    offsets[1:] = np.cumsum(padded[:-1])
    if (lengths & 1).any():
        out = np.zeros(int(padded.sum()) + 2, dtype=np.float64)
        src = np.concatenate(([0], np.cumsum(lengths.astype(np.int64))))
        shift = offsets - src[:-1]
        dst_idx = np.arange(sig.shape[0], dtype=np.int64) + np.repeat(shift, lengths)
        out[dst_idx] = sig
    else:
        out = np.concatenate((sig, np.zeros(2)))
    return out, offsets, lengths, rev, truth
