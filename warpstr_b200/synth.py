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


# ---------------------------------------------------------------------------------------------------
# Multi-locus panels: many loci, a fixed number of reads each, sharded over the GPUs
# of a box.  A locus's reads are `base_reads` noiseless (allele, dwell) realisations, each used
# reads_per_locus / base_reads times with its own Gaussian noise.  The noise is drawn on the GPU from a
# generator seeded per locus and laid out over ALL reads of the locus in read order, so a read's samples
# do not depend on how the panel is sharded: N ranks and one rank see bit-identical reads.
# ---------------------------------------------------------------------------------------------------
PANEL_MIX = ('AAAT', 'HD', 'FMR1', 'FMR1_MGG', 'DM2', 'CAN', 'RFC1', 'C9ORF72_100', 'HD', 'FMR1')


def make_panel(n_loci: int, seed: int = 0, patterns: Tuple[str, ...] = PANEL_MIX) -> List[SynthLocus]:
    """``n_loci`` loci cycling through ``patterns``, each with its own random flanks."""
    return [make_locus(f'{patterns[i % len(patterns)]}_{i}', seed=seed * 1000 + i, recipe=patterns[i % len(patterns)])
            for i in range(n_loci)]


def _panel_base_one(args):
    locus, base_reads, seed = args
    return make_read_batch(locus, base_reads, seed=seed, noise=0.0)


def make_panel_base(loci: List[SynthLocus], base_reads: int, seed: int = 0, workers: int = 1):
    """Noiseless base reads of every locus: list of (signal, offsets, lengths, reverse, truth_len).
    ``workers`` > 1 spreads the loci over forked processes (call it before CUDA is initialised)."""
    jobs = [(loc, base_reads, seed * 7919 + i) for i, loc in enumerate(loci)]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(min(workers, len(jobs))) as pool:
            return pool.map(_panel_base_one, jobs, chunksize=1)
    return [_panel_base_one(j) for j in jobs]


def panel_read_table(base, reads_per_locus: int):
    """Per read of the panel (global id = locus * reads_per_locus + k): locus, base read, length, strand."""
    n_loci = len(base)
    B = len(base[0][2])
    k = np.arange(reads_per_locus, dtype=np.int64)
    bidx = np.tile(k % B, n_loci)
    locus = np.repeat(np.arange(n_loci, dtype=np.int64), reads_per_locus)
    lengths = np.concatenate([b[2][k % B] for b in base]).astype(np.int32)
    rev = np.concatenate([b[3][k % B] for b in base]).astype(np.uint8)
    truth = np.concatenate([b[4][k % B] for b in base]).astype(np.int32)
    return locus, bidx, lengths, rev, truth


def materialize_panel_reads(base, reads_per_locus: int, ids: np.ndarray, noise: float, seed: int, device):
    """Signals of the panel reads ``ids`` (ascending global ids) on ``device``: float64, concatenated with
    even starts.  Returns (d_sig, offsets int64[n], lengths int32[n])."""
    import torch
    ids = np.asarray(ids, dtype=np.int64)
    n_loci = len(base)
    B = len(base[0][2])
    loc_of = ids // reads_per_locus
    k_of = ids % reads_per_locus
    lengths = np.empty(len(ids), dtype=np.int32)
    for L in range(n_loci):
        sel = loc_of == L
        if sel.any():
            lengths[sel] = base[L][2][k_of[sel] % B]
    padded = (lengths.astype(np.int64) + 1) & ~1
    offsets = np.zeros(len(ids), dtype=np.int64)
    offsets[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum()) + 2
    d_sig = torch.zeros(total, dtype=torch.float64, device=device)
    kall = np.arange(reads_per_locus, dtype=np.int64)
    for L in range(n_loci):
        sel = np.flatnonzero(loc_of == L)
        if not len(sel):
            continue
        bsig, boff, blen = base[L][0], base[L][1], base[L][2]
        all_len = blen[kall % B].astype(np.int64)              # every read of the locus, in read order
        cum = np.zeros(reads_per_locus + 1, dtype=np.int64)
        cum[1:] = np.cumsum(all_len)
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed) * 1000003 + L)
        z = torch.randn(int(cum[-1]), generator=gen, dtype=torch.float64, device=device)
        d_base = torch.from_numpy(bsig).to(device)
        ks = k_of[sel]
        ln = torch.from_numpy(all_len[ks]).to(device)
        rep = torch.repeat_interleave(torch.arange(len(sel), device=device), ln)
        first = torch.cumsum(ln, 0) - ln
        t = torch.arange(int(ln.sum().item()), device=device) - first[rep]
        src = torch.from_numpy(boff[ks % B]).to(device)[rep] + t
        nz = torch.from_numpy(cum[ks]).to(device)[rep] + t
        dst = torch.from_numpy(offsets[sel]).to(device)[rep] + t
        d_sig[dst] = d_base[src] + noise * z[nz]
        del z, d_base, rep, t, src, nz, dst
    return d_sig, offsets, lengths
